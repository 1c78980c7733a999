// nxgpu_job.cu — interprets one NX job descriptor (nx_gzip_crb_cpb_t, reference
// inc_nx/nxu.h:286-616) and runs it on the GPU.  Stub: filled in after the batch path.
#include <errno.h>
#include <stdint.h>
#include "common.cuh"
#include "../../include/nxgpu.h"
namespace nxgpu {
int run_job_impl(nxgpu_ctx *ctx, uint8_t *crb_cpb) { (void)ctx; (void)crb_cpb; return -EAGAIN; }
}
