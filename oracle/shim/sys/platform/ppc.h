/* x86 stand-in for glibc's POWER-only <sys/platform/ppc.h>, so the reference's
 * host code (inc_nx/nxu.h:63, lib/nx_zlib.h:56) compiles in place for oracle/_ref.
 * Test infrastructure only. */
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
