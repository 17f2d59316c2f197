// tsc_scan_launch.h — what tsc_scan.cu (geometry, parameters) and the per-metric kernel
// translation units (tsc_scan_l2 / _ip / _cos.cu, compiled in parallel) share.
#pragma once

#include "tsc_index.h"
#include "tsc_scan.cuh"

namespace tsc {

struct ScanPlan {
  int warps, rows, stages, qb;
  uint32_t stage_bytes, sort_cap;
  size_t smem;
};

int32_t scan_dispatch_l2(Index *ix, const ScanParams &p, const ScanPlan &pl, bool sparse, cudaStream_t st);
int32_t scan_dispatch_ip(Index *ix, const ScanParams &p, const ScanPlan &pl, bool sparse, cudaStream_t st);
int32_t scan_dispatch_cos(Index *ix, const ScanParams &p, const ScanPlan &pl, bool sparse, cudaStream_t st);

}  // namespace tsc
