from .cache_value import (get_node_heat, get_cache_nids_by_degree, get_structure_space,
                          get_feature_space, get_available_memory)
