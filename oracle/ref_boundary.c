/*
 * ref_boundary.c — the six boundary symbols (lib/nx_zlib.h:625-629, inc_nx/nxu.h:71,
 * lib/crc32_ppc.c:30) for the x86 build of the UNMODIFIED reference host code in
 * oracle/_ref/libnxz_ref.so.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * nxu_run_job is served by the software engine in nxemu.c.
 */
#include <errno.h>
#include <stdint.h>
#include <time.h>
#include "oracle.h"

struct nx_dev_t;
struct nx_gzip_crb_cpb_t;
int oracle_nxemu_run_job(void *crb_cpb);

uint64_t tb_freq;

struct dev_prefix { int i[7]; int pid; void *paste_addr; int fd; int function; };

int nx_function_begin(int function, int pri, struct nx_dev_t *h)
{
	struct dev_prefix *d = (struct dev_prefix *)h;
	(void)pri;
	if (function != 2) { errno = EINVAL; return -1; }
	d->paste_addr = (void *)d;     /* any non-NULL value: lib/gzip_vas.c:294 */
	d->fd = -1;
	d->function = function;
	return 0;
}

int nx_function_end(struct nx_dev_t *h) { (void)h; return 0; }

uint64_t nx_wait_ticks(uint64_t ticks, uint64_t acc, int do_sleep)
{
	struct timespec ts = { 0, (long)(ticks * 1000ull / 512ull) };
	(void)do_sleep;
	nanosleep(&ts, NULL);
	return acc + ticks;
}

int nxu_run_job(struct nx_gzip_crb_cpb_t *c, struct nx_dev_t *h)
{
	(void)h;
	return oracle_nxemu_run_job(c);
}

unsigned int __crc32_vpmsum(unsigned int crc, const void *p, unsigned long len)
{
	return oracle_crc32_raw(crc, p, len);
}
