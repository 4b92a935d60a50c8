// C-ABI plumbing: error text, launch accounting, dstag_gemm dispatch.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace dstag {

static thread_local char g_err[512] = {0};
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int gemm_simt(const dstag_gemm_t* g, cudaStream_t stream);
int gemm_tc(const dstag_gemm_t* g, cudaStream_t stream);
bool gemm_tc_supported(const dstag_gemm_t* g);

static int validate(const dstag_gemm_t* g) {
  DSTAG_CHECK_ARG(g != nullptr, "dstag_gemm: null descriptor");
  DSTAG_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0 && g->batch > 0,
                  "dstag_gemm: empty problem M=%d N=%d K=%d batch=%d", g->M, g->N, g->K, g->batch);
  DSTAG_CHECK_ARG(g->A.ptr && g->B.ptr && g->C.ptr, "dstag_gemm: null A/B/C pointer");
  DSTAG_CHECK_ARG(g->split_k >= 1, "dstag_gemm: split_k must be >= 1");
  DSTAG_CHECK_ARG(g->accumulate >= DSTAG_ACC_STORE && g->accumulate <= DSTAG_ACC_ATOMIC,
                  "dstag_gemm: bad accumulate mode %d", g->accumulate);
  DSTAG_CHECK_ARG(g->split_k == 1 || (!g->relu && (g->accumulate == DSTAG_ACC_ATOMIC ||
                                                   (g->accumulate == DSTAG_ACC_STORE && g->split_stride != 0))),
                  "dstag_gemm: split_k > 1 needs atomic accumulation or per-split slices (split_stride), and no relu");
  DSTAG_CHECK_ARG(g->split_stride == 0 || (g->accumulate == DSTAG_ACC_STORE && !g->Cmul.ptr),
                  "dstag_gemm: split_stride needs accumulate == STORE and no Cmul");
  const dstag_axis_t* ax[] = {&g->A.batch, &g->A.row, &g->A.col, &g->B.batch, &g->B.row, &g->B.col,
                              &g->C.batch, &g->C.row, &g->C.col};
  for (const dstag_axis_t* a : ax) DSTAG_CHECK_ARG(a->d1 >= 1, "dstag_gemm: axis d1 must be >= 1");
  return 0;
}

}  // namespace dstag

extern "C" {

int dstag_abi_version(void) { return 19; }
const char* dstag_last_error(void) { return dstag::g_err; }
int64_t dstag_launch_count(void) { return dstag::g_launches.load(); }
void dstag_reset_launch_count(void) { dstag::g_launches.store(0); }

int dstag_gemm_uses_tensor_cores(const dstag_gemm_t* g) {
  if (dstag::validate(g) != 0) return 0;
  return (g->precision != DSTAG_PREC_FP32 && dstag::gemm_tc_supported(g)) ? 1 : 0;
}

int dstag_gemm(const dstag_gemm_t* g, void* stream) {
  const int rc = dstag::validate(g);
  if (rc != 0) return rc;
  if (g->precision != DSTAG_PREC_FP32 && dstag::gemm_tc_supported(g))
    return dstag::gemm_tc(g, (cudaStream_t)stream);
  return dstag::gemm_simt(g, (cudaStream_t)stream);
}

}  // extern "C"
