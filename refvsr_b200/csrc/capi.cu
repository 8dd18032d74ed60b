// C-ABI glue: error reporting, launch accounting and the rv_conv2d dispatcher.
#include <cstdarg>

#include "common.cuh"

namespace rv {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int conv2d_simt(const rv_conv_desc* d, cudaStream_t st);
int conv2d_tc(const rv_conv_desc* d, cudaStream_t st);
int conv2d_tc_plan(const rv_conv_desc* d, int max_smem, int num_sms, int* out);
int set_conv_cta_cap(int cap);

}  // namespace rv

extern "C" const char* rv_last_error(void) { return rv::g_last_error.c_str(); }
extern "C" int rv_version(void) { return 100; }
extern "C" uint64_t rv_launch_count(void) { return rv::g_launches.load(); }

extern "C" int rv_conv2d(const rv_conv_desc* d, void* stream) {
  RV_REQUIRE(d != nullptr, "rv_conv2d: null descriptor");
  RV_REQUIRE(d->src0 && d->wpack && d->bias && d->out, "rv_conv2d: null src0/wpack/bias/out");
  RV_REQUIRE(d->H > 0 && d->W > 0 && d->c0 > 0 && d->cout > 0 && d->kh > 0 && d->kw > 0 && d->stride > 0,
             "rv_conv2d: bad geometry H=%d W=%d c0=%d cout=%d k=%dx%d stride=%d", d->H, d->W, d->c0,
             d->cout, d->kh, d->kw, d->stride);
  if (d->impl == RV_CONV_IMPL_TC) return rv::conv2d_tc(d, (cudaStream_t)stream);
  if (d->impl == RV_CONV_IMPL_SIMT) return rv::conv2d_simt(d, (cudaStream_t)stream);
  return rv::fail(RV_E_INVALID, "rv_conv2d: unknown impl %d", d->impl);
}

extern "C" int rv_set_conv_cta_cap(int cap) { return rv::set_conv_cta_cap(cap); }

extern "C" int rv_conv2d_tc_plan(const rv_conv_desc* d, int max_smem_optin, int num_sms, int32_t* out8) {
  RV_REQUIRE(d != nullptr && out8 != nullptr && max_smem_optin > 0 && num_sms > 0, "rv_conv2d_tc_plan: bad arguments");
  return rv::conv2d_tc_plan(d, max_smem_optin, num_sms, out8);
}
