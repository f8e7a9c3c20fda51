"""inference/inference_tabledetection.py:40-57 of the reference: load_model(model_path) -> (table_model, transform_table)."""
from upliftingtabletennis_b200.interface import load_table_model as load_model  # noqa: F401
