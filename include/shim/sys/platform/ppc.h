/* x86 stand-in for glibc's POWER-only <sys/platform/ppc.h>: the one header the reference's host
 * code (inc_nx/nxu.h:63, lib/nx_zlib.h:56) needs to compile unmodified on the B200 box's host
 * (INTEGRATION.md).  The timebase is the 512 MHz tick the library's wait loops are written for. */
#ifndef NXGPU_SHIM_PPC_H
#define NXGPU_SHIM_PPC_H
#include <stdint.h>
#include <time.h>
static inline uint64_t __ppc_get_timebase(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 512000000ull + (uint64_t)ts.tv_nsec * 512ull / 1000ull;
}
static inline uint64_t __ppc_get_timebase_freq(void) { return 512000000ull; }
#endif
