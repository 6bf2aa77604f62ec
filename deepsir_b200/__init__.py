"""deepsir_b200 — B200-native (sm_100a) drop-in for DeepSIR's correspondence-and-pose hot path.

Host-side mirror of the reference's operator interface (same names and argument meaning as
network/matchnet.py, network/model.py, network/tools.py, common/math/se3_torch.py and the
torch_points_kernels.knn contract) over the C-ABI library libdeepsir_b200.so.  CUDA only.
"""
from ._lib import DeepSIRError, build, lib, LIB_PATH, EXPORTS  # noqa: F401
from ._lib import KNN_AUTO, KNN_BRUTE, KNN_GRID, KNN_TREE, MATCH_AUTO, MATCH_FP32, MATCH_TC  # noqa: F401
from .match import (square_distance, square_distance_V2, match_features, match_features_V2, feat_dist,  # noqa: F401
                    match_argmin, match_soft, sinkhorn_implicit, log_optimal_transport_implicit, compute_affinity,
                    gather_neighbour_V3)
from .kabsch import (compute_rigid_transform, compute_rigid_transform_2, kabsch_gather, kabsch_moments,  # noqa: F401
                     kabsch_from_moments, kabsch_soft)
from .knn import knn, nn_search, nn_search_cloud, nn_search_pair  # noqa: F401
from .loop import align_loop, pred_pairs  # noqa: F401
from .pipeline import RegistrationPipeline  # noqa: F401
from .graphs import GraphedRegistration  # noqa: F401
from .graph import (gather_neighbour, gather_neighbour_V2, gather_neighbour_V4, relative_pos_encoding, random_sample,  # noqa: F401
                    nearest_interpolation, sinkhorn, log_optimal_transport)
from .keypoint import score_fun, feat_score, topk  # noqa: F401
from . import metrics  # noqa: F401
from . import se3 as se3_torch  # noqa: F401
from . import se3, synth  # noqa: F401

__version__ = "0.1.0"
