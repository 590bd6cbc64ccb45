"""skimage stub so the reference's ``model.py`` imports in the build container
(skimage is absent; only metrics.py needs it and the golden
vectors never call those metrics)."""
import sys
import types

if "skimage" not in sys.modules:
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.metrics")
    skm.structural_similarity = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError)
    skm.peak_signal_noise_ratio = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError)
    sk.metrics = skm
    sys.modules["skimage"] = sk
    sys.modules["skimage.metrics"] = skm
