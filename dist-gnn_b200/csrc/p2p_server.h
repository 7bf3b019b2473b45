// p2p_server.h - internal layout of dgs_p2p_server (C-ABI handle in include/dgs_b200.h).
#pragma once
#include <stdint.h>

#include "../../include/dgs_b200.h"

struct dgs_p2p_server {
  int world;
  int rank;
  int owns_local;      // local shard was allocated by us
  int ipc_opened;      // peers were opened with cudaIpcOpenMemHandle (legacy path)
  int vmm;             // shards are cuMemCreate allocations mapped with cuMemMap (fd exchange)
  unsigned long long vmm_handle[DGS_MAX_DEVICES];  // CUmemGenericAllocationHandle per mapped shard
  int64_t vmm_size[DGS_MAX_DEVICES];               // mapped (granularity-rounded) size
  void *ptrs[DGS_MAX_DEVICES];
  int64_t nbytes[DGS_MAX_DEVICES];
};
