from .cache_value import *
from .cache_value import (B200_COST_MODEL, get_node_heat, get_cache_nids_by_degree, get_hot_nids_local,
                          get_hot_nids_p2p_global, get_structure_space, get_feature_space,
                          get_node_value, get_cache_nids_local, get_cache_nids_selfish,
                          get_cache_nids_selfless, compute_total_value_selfish,
                          compute_total_value_selfless, choose_cache_policy, get_available_memory)
