import _bootstrap  # noqa: F401
from pcfm.functional import (avg_voxelize, ball_query, furthest_point_sample, gather, grouping, huber_loss, kl_loss,
                             logits_mask, nearest_neighbor_interpolate, trilinear_devoxelize)
