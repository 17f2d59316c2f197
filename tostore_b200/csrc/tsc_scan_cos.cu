// tsc_scan_cos.cu — K1 / K6 kernels for one metric (see tsc_scan_metric.inc)
#define TSC_SCAN_METRIC kCos
#define TSC_SCAN_FN scan_dispatch_cos
#include "tsc_scan_metric.inc"
