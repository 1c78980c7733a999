/* placeholder replaced below */
#include <errno.h>
int oracle_nxemu_run_job(void *c) { (void)c; return -EAGAIN; }
