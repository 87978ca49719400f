// Error plumbing, version and the host-side twiddle table builder.
#include <cmath>
#include <cstdarg>

#include "pdes_common.cuh"

namespace pdes {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PDES_ERR_LAUNCH;
  }
  return PDES_OK;
}

void* tensor_map_encoder() {
#ifdef PDES_CPU_EMU
  return nullptr;
#else
  static void* fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = sym;
  }
  return fn;
#endif
}

}  // namespace pdes

extern "C" {

int pdes_version(void) { return 100; }

const char* pdes_last_error(void) { return pdes::g_err; }

int pdes_is_cuda_build(void) {
#ifdef PDES_CPU_EMU
  return 0;
#else
  return 1;
#endif
}

size_t pdes_tables_floats(int H, int W, int m1, int m2) {
  if (H <= 0 || W <= 0 || m1 <= 0 || m2 <= 0) return 0;
  return pdes::table_layout(H, W, m1, m2).total;
}

int pdes_tables_fill(int H, int W, int m1, int m2, float* buf) {
  PDES_REQUIRE(buf != nullptr, PDES_ERR_ARG, "pdes_tables_fill: null buffer");
  PDES_REQUIRE(H > 0 && W > 0 && m1 > 0 && m2 > 0, PDES_ERR_ARG, "pdes_tables_fill: non-positive size");
  PDES_REQUIRE(m1 <= H && m2 <= W / 2 + 1, PDES_ERR_ARG,
               "modes (%d,%d) exceed the grid (%d,%d): need m1 <= H and m2 <= W/2+1", m1, m2, H, W);
  const pdes::TableLayout t = pdes::table_layout(H, W, m1, m2);
  for (size_t i = 0; i < t.total; ++i) buf[i] = 0.0f;
  const double two_pi = 6.283185307179586476925286766559;
  for (int j = 0; j < H; ++j) {
    const double a = two_pi * (double)j / (double)H;
    buf[t.twh + 2 * j] = (float)std::cos(a);
    buf[t.twh + 2 * j + 1] = (float)std::sin(a);
  }
  for (int j = 0; j <= m1; ++j)
    for (int pp = 0; pp < t.npp; ++pp) {
      const long r = ((long)j * (long)(pp <= H / 2 ? pp : 0)) % (long)H;
      const double a = two_pi * (double)r / (double)H;
      buf[t.twp + ((size_t)j * t.npp + pp) * 2] = (float)std::cos(a);
      buf[t.twp + ((size_t)j * t.npp + pp) * 2 + 1] = (float)std::sin(a);
    }
  for (int l = 0; l < m2; ++l)
    for (int w = 0; w < W; ++w) {
      const long r = ((long)l * (long)w) % (long)W;
      const double a = two_pi * (double)r / (double)W;
      buf[t.tw2 + ((size_t)l * (W + 1) + w) * 2] = (float)std::cos(a);
      buf[t.tw2 + ((size_t)l * (W + 1) + w) * 2 + 1] = (float)std::sin(a);
    }
  for (int l = 0; l < m2; ++l) {
    double c = 2.0;
    if (l == 0) c = 1.0;
    if (W % 2 == 0 && l == W / 2) c = 1.0;
    const double s = c / ((double)H * (double)W);
    buf[t.herm + l] = (float)s;
    for (int w = 0; w < W; ++w) {
      // reduce l*w mod W in integers so the angle stays small and exact
      const long r = ((long)l * (long)w) % (long)W;
      const double a = two_pi * (double)r / (double)W;
      const double cs = std::cos(a), sn = std::sin(a);
      buf[t.twa + (size_t)w * t.nc4 + 2 * l] = (float)cs;
      buf[t.twa + (size_t)w * t.nc4 + 2 * l + 1] = (float)(-sn);
      buf[t.tinv_f + (size_t)(2 * l) * W + w] = (float)(s * cs);
      buf[t.tinv_f + (size_t)(2 * l + 1) * W + w] = (float)(-s * sn);
      buf[t.tinv_b + (size_t)(2 * l) * W + w] = (float)cs;
      buf[t.tinv_b + (size_t)(2 * l + 1) * W + w] = (float)(-sn);
    }
  }
  return PDES_OK;
}

}  // extern "C"
