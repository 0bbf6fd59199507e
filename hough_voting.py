"""`import hough_voting` -- op names used by BASELINE.json's north_star (vote / back_project)."""
from canonicalvoting_b200.hough_voting import (HoughVoting, HVFunction, back_project, back_project_numpy,  # noqa: F401
                                                 get_iou_obb, iou_matrix, nms_per_class, vote, vote_host)
from canonicalvoting_b200.proposals import HoughVotingModule, project_y  # noqa: F401,E402  (sunrgbd/brnetcanon.py:104-162)
