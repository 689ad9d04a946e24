// perm_host.cu — host-side index-vector form of the reference's hierarchical
// swap-permutation sampler (run.py:107-182 my_swap_h/w + block_permutation,
// driven as in run.py:436-507).  The reference builds dense 0/1 matrices in
// Python loops (7 ms per 96x96 matrix); here one matrix is a few
// hundred integer swaps on an index vector.  The caller pre-draws the uniforms
// from the SAME numpy MT19937 stream (np.random.uniform(size=n) yields the same
// sequence as n scalar calls), so the integer tile-index grid stays bit-exact.
#include "common.cuh"

// One "swapped identity": two sweeps (0..n-1, n-1..0), one uniform per visit.
static void swap_sweep(const double* u, int n, int32_t* m) {
  for (int i = 0; i < n; ++i) m[i] = i;
  if (n <= 1) return;
  int q = 0;
  for (int sweep = 0; sweep < 2; ++sweep) {
    for (int s = 0; s < n; ++s) {
      const int i = sweep == 0 ? s : n - 1 - s;
      const double p = u[q++];
      int j = -1;
      if (i == 0) {
        if (p < 0.5) j = 1;
      } else if (i == n - 1) {
        if (p < 0.5) j = n - 2;
      } else if (p < 1.0 / 3.0) {
        j = i + 1;
      } else if (p > 2.0 / 3.0) {
        j = i - 1;
      }
      if (j >= 0) {
        int32_t tmp = m[i];
        m[i] = m[j];
        m[j] = tmp;
      }
    }
  }
}

// count index vectors of `length`, `levels` hierarchy levels (block size 2^l).
// Row map of P_h (P_h <- P_h @ kron(swap_l, I)) and column map of P_w
// (P_w <- kron(swap_l, I) @ P_w) obey the same recurrence r <- b_l[r].
extern "C" int tmx_perm_indices_from_uniforms(const double* u, int64_t n_u, int length, int levels, int count,
                                              int32_t* out, int64_t* consumed) {
  TMX_REQUIRE(u && out && consumed, TMX_ERR_ARG, "tmx_perm_indices_from_uniforms: NULL argument");
  TMX_REQUIRE(length > 0 && levels >= 0 && levels < 31 && count >= 0, TMX_ERR_SHAPE,
              "tmx_perm_indices_from_uniforms: bad length=%d levels=%d count=%d", length, levels, count);
  for (int l = 0; l < levels; ++l)
    TMX_REQUIRE(length % (1 << l) == 0, TMX_ERR_SHAPE,
                "tmx_perm_indices_from_uniforms: length %d not divisible by block size %d", length, 1 << l);
  int64_t need = 0;
  for (int l = 0; l < levels; ++l) {
    int n = length >> l;
    if (n > 1) need += 2 * (int64_t)n;
  }
  TMX_REQUIRE(need * count <= n_u, TMX_ERR_SHAPE,
              "tmx_perm_indices_from_uniforms: %lld uniforms needed, %lld given", (long long)(need * count),
              (long long)n_u);
  int32_t* m = new int32_t[length];
  int32_t* b = new int32_t[length];
  int64_t q = 0;
  for (int c = 0; c < count; ++c) {
    int32_t* r = out + (int64_t)c * length;
    for (int i = 0; i < length; ++i) r[i] = i;
    for (int l = 0; l < levels; ++l) {
      const int bs = 1 << l, n = length >> l;
      swap_sweep(u + q, n, m);
      if (n > 1) q += 2 * (int64_t)n;
      for (int i = 0; i < n; ++i)
        for (int t = 0; t < bs; ++t) b[i * bs + t] = m[i] * bs + t;  // kron(perm, I_bs)
      for (int i = 0; i < length; ++i) r[i] = b[r[i]];
    }
  }
  delete[] m;
  delete[] b;
  *consumed = q;
  return TMX_OK;
}
