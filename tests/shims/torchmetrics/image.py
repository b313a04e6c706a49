import importlib

_m = importlib.import_module("vae-channel-dynamics_b200.metrics")
PeakSignalNoiseRatio = _m.PeakSignalNoiseRatio
StructuralSimilarityIndexMeasure = _m.StructuralSimilarityIndexMeasure
