"""inference/inference_balldetection.py:40-61 of the reference: load_model(model_path) -> (ball_model, transform_ball)."""
from upliftingtabletennis_b200.interface import load_ball_model as load_model  # noqa: F401
