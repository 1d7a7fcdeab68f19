// Reference-side shim (checker only): exposes the reference CLI's compute_error_metrics() - which prints its results -
// as a C symbol, so tests can run the UNMODIFIED astcenccli_error_metrics.cpp (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libastcenc_ref_metrics.so) and parse what it prints.
#include "astcenccli_internal.h"

extern "C" __attribute__((visibility("default")))
void ref_compute_error_metrics(int hdr, int normal, int input_components, const astcenc_image* img1, const astcenc_image* img2,
                               int fstop_lo, int fstop_hi) {
	compute_error_metrics(hdr != 0, normal != 0, input_components, img1, img2, fstop_lo, fstop_hi);
	fflush(stdout);
}
