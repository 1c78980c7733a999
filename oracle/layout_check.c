/* layout_check.c — proves include/nxgpu.h's descriptor offsets and handle prefix
 * equal the reference's (inc_nx/nxu.h:286-616, lib/nx_zlib.h:178-194).
 * TEST INFRASTRUCTURE ONLY; compiled only where /root/reference exists. */
#include <stdio.h>
#include <stddef.h>
#include "nx_zlib.h"
#define NXGPU_NO_DROPIN_DECLS
#include "nxgpu.h"
#define CK(a, b) do { if ((size_t)(a) != (size_t)(b)) { printf("MISMATCH %s=%zu %s=%zu\n", #a, (size_t)(a), #b, (size_t)(b)); bad = 1; } } while (0)
int main(void)
{
	int bad = 0;
	CK(sizeof(nx_gzip_crb_cpb_t), NXGPU_CRB_CPB_SIZE);
	CK(offsetof(nx_gzip_crb_cpb_t, crb.gzip_fc), NXGPU_CRB_FC);
	CK(offsetof(nx_gzip_crb_cpb_t, crb.csb_address), NXGPU_CRB_CSB_ADDR);
	CK(offsetof(nx_gzip_crb_cpb_t, crb.source_dde), NXGPU_CRB_SRC_DDE);
	CK(offsetof(nx_gzip_crb_cpb_t, crb.target_dde), NXGPU_CRB_DST_DDE);
	CK(offsetof(nx_gzip_crb_cpb_t, crb.csb), NXGPU_CRB_CSB);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb), NXGPU_CPB);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.in_adler), NXGPU_CPB_IN_ADLER);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.in_crc), NXGPU_CPB_IN_CRC);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.in_histlen), NXGPU_CPB_IN_HISTLEN);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.in_sfbt), NXGPU_CPB_IN_SFBT);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.in_dht), NXGPU_CPB_IN_DHT);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_adler), NXGPU_CPB_OUT_ADLER);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_crc), NXGPU_CPB_OUT_CRC);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_tebc), NXGPU_CPB_OUT_TEBC);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_sfbt), NXGPU_CPB_OUT_SFBT);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_spbc_comp), NXGPU_CPB_OUT_SPBC_COMP);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_lzcount), NXGPU_CPB_OUT_LZCOUNT);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_dht), NXGPU_CPB_OUT_DHT);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_spbc_decomp), NXGPU_CPB_OUT_SPBC_DECOMP);
	CK(offsetof(nx_gzip_crb_cpb_t, cpb.out_spbc_comp_with_count), NXGPU_CPB_OUT_SPBC_COMP_WITH_COUNT);
	CK(offsetof(struct nx_dev_t, paste_addr), offsetof(struct nxgpu_dev_prefix, paste_addr));
	CK(offsetof(struct nx_dev_t, fd), offsetof(struct nxgpu_dev_prefix, fd));
	CK(offsetof(struct nx_dev_t, function), offsetof(struct nxgpu_dev_prefix, function));
	CK(offsetof(struct nx_dev_t, creator_pid), offsetof(struct nxgpu_dev_prefix, creator_pid));
	printf(bad ? "layout_check FAILED\n" : "layout_check ok\n");
	return bad;
}
