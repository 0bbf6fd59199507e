"""`import hough_voting` -- op names used by BASELINE.json's north_star (vote / back_project)."""
from canonicalvoting_b200.hough_voting import (HoughVoting, HVFunction, back_project, back_project_numpy,  # noqa: F401
                                                 get_iou_obb, iou_matrix, nms_per_class, vote, vote_host)
