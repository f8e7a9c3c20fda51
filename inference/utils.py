"""inference/utils.py of the reference: the per-trajectory stage functions, on libttk."""
from upliftingtabletennis_b200.interface import (HEIGHT, WIDTH, BALL_VISIBLE, KEYPOINT_VISIBLE, KEYPOINT_INVISIBLE,  # noqa: F401
                                                  _uplifting_transform, calibrate_camera, filter_trajectory_ball, filter_trajectory_table)
from upliftingtabletennis_b200.trajectory import (extract_position_ball, extract_position_table, process_trajectory_ball,  # noqa: F401
                                                   process_trajectory_table, process_trajectory_uplifting)
