"""inference/inference_uplifting.py:33-58 of the reference: load_model(model_path) -> (uplifting_model, transform, transform_mode)."""
from upliftingtabletennis_b200.interface import load_uplifting_model as load_model  # noqa: F401
