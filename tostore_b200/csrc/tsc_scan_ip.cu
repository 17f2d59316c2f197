// tsc_scan_ip.cu — K1 / K6 kernels for one metric (see tsc_scan_metric.inc)
#define TSC_SCAN_METRIC kIP
#define TSC_SCAN_FN scan_dispatch_ip
#include "tsc_scan_metric.inc"
