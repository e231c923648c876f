// K15  CaNet dense-comparison input.
//
// replaces networks/canet.py:172-180
//   z   = mean over shots of  sum_x f m / (sum_x m + 1e-5)        (K1: pemp_map_pool_lowres, foreground only)
//   out = cat(qry_fts, z tiled to [BQ, c, h, w])                  [BQ, 2c, h, w]
// The reference materialises the tiled z ([BQ, c, h, w]) and then copies both halves again in torch.cat; here one kernel
// reads the query features in place (episode stride) and writes the 2c-channel tensor once:
// algorithmic bytes = Q*c*hw*4 read + 2*Q*c*hw*4 written per episode.
#include "common.cuh"

namespace {

// one warp per output row (n, ch): coalesced copy of the query row or broadcast of z[b][ch - c]
__global__ void __launch_bounds__(256)
canet_concat_kernel(const float* __restrict__ qry, long long ep_stride, int Q, const float* __restrict__ z, int c, int hw,
                    long long rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    const long long n = r / (2 * c);
    const int ch = static_cast<int>(r - n * 2 * c);
    const long long b = n / Q, qi = n - b * Q;
    float* dst = out + r * hw;
    if (ch < c) {
      const float* src = qry + b * ep_stride + (qi * c + ch) * static_cast<long long>(hw);
      for (int i = lane; i < hw; i += 32) dst[i] = __ldg(src + i);
    } else {
      const float v = __ldg(z + b * c + (ch - c));
      for (int i = lane; i < hw; i += 32) dst[i] = v;
    }
  }
}

}  // namespace

extern "C" int pemp_canet_concat(const float* qry, long long qry_episode_stride, const float* z, int N, int Bp, int c, int hw,
                                 float* out, pemp_stream_t stream) {
  PEMP_REQUIRE(qry && z && out, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && Bp > 0 && N % Bp == 0 && c > 0 && hw > 0, PEMP_E_SHAPE);
  const int Q = N / Bp;
  const long long rows = static_cast<long long>(N) * 2 * c;
  const long long ep = qry_episode_stride ? qry_episode_stride : static_cast<long long>(Q) * c * hw;
  const unsigned blocks = static_cast<unsigned>(llmin((rows + 7) / 8, 148LL * 16));
  canet_concat_kernel<<<blocks, 256, 0, as_stream(stream)>>>(qry, ep, Q, z, c, hw, rows, out);
  return launch_status();
}
