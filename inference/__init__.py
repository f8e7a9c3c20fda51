"""Import shim: the reference's ``inference`` package paths (``inference.utils``, ``inference.inference_balldetection.load_model``,
``inference.inference_tabledetection.load_model``, ``inference.inference_uplifting.load_model``) backed by
``upliftingtabletennis_b200``.  Only the functions on the inference hot path exist; the evaluation loops over the TTHQ / TTST / TT3D
datasets are out of scope."""
