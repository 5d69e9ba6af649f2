// G1 MSM entry points - placeholder until the Pippenger kernels land (returns JB_ERR_UNSUPPORTED).
#include "ctx.hpp"

void jb_ctx::msm_release() {}

extern "C" {
int jb_srs_upload_affine(jb_ctx* c, const uint64_t*, size_t, jb_srs*) { return c ? c->fail(JB_ERR_UNSUPPORTED, "msm not built") : JB_ERR_INVALID; }
int jb_srs_upload_jacobian(jb_ctx* c, const uint64_t*, size_t, jb_srs*) { return c ? c->fail(JB_ERR_UNSUPPORTED, "msm not built") : JB_ERR_INVALID; }
int jb_srs_len(jb_ctx* c, jb_srs, size_t*) { return c ? c->fail(JB_ERR_UNSUPPORTED, "msm not built") : JB_ERR_INVALID; }
int jb_srs_download_affine(jb_ctx* c, jb_srs, uint64_t*, size_t) { return c ? c->fail(JB_ERR_UNSUPPORTED, "msm not built") : JB_ERR_INVALID; }
int jb_srs_free(jb_ctx* c, jb_srs) { return c ? c->fail(JB_ERR_UNSUPPORTED, "msm not built") : JB_ERR_INVALID; }
int jb_msm_g1(jb_ctx* c, jb_srs, size_t, const uint64_t*, size_t, uint64_t*) { return c ? c->fail(JB_ERR_UNSUPPORTED, "msm not built") : JB_ERR_INVALID; }
int jb_msm_g1_table(jb_ctx* c, jb_srs, size_t, jb_table, size_t, uint64_t*) { return c ? c->fail(JB_ERR_UNSUPPORTED, "msm not built") : JB_ERR_INVALID; }
}
