// tsc_scan_l2.cu — K1 / K6 kernels for one metric (see tsc_scan_metric.inc)
#define TSC_SCAN_METRIC kL2
#define TSC_SCAN_FN scan_dispatch_l2
#include "tsc_scan_metric.inc"
