/*
 * tzk_oracle.c — C/OpenMP restatement of the hot path's heavy steps, for the CPU baseline.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/tzk_oracle.py for the rules): used by tests/ as a second checker
 * (it is itself checked against the numpy oracle in tests/test_oracle_c.py) and by bench.py as the timed
 * "reference CPU path" (cpu_baseline.kind = "port").  It mirrors what the reference executes on host cores
 * through [EXT] fbgemm-gpu's CPU TBE kernels and ATen: same arithmetic as tzk_oracle.py, multi-threaded.
 *
 *   orc_pooled_lookup   <- tzrec/modules/embedding.py:930  ([EXT] TBE forward, App. A.3)
 *   orc_fused_update    <- tzrec/main.py:774-781           ([EXT] TBE fused backward+optimizer EXACT, App. A.10)
 *   orc_dot_interact_*  <- tzrec/modules/interaction.py:80-91 + tzrec/models/dlrm.py:113-131
 *   orc_fm_*            <- tzrec/modules/fm.py:28-42
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define POOL_MEAN 1
#define OPT_SGD 0
#define OPT_ADAGRAD 1
#define OPT_ROWWISE_ADAGRAD 2

int orc_num_threads(void) { return omp_get_max_threads(); }

/* out[b, col_f : +D_f] = pool over bag (f,b), sequential fp32 adds in list order */
void orc_pooled_lookup(const float* weights, const int64_t* w_off, const int64_t* rows, const int32_t* dim,
                       const int32_t* col, const int32_t* pool, const int64_t* ids, const int64_t* offsets,
                       int32_t F, int32_t B, float* out, int64_t ld_out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int f = 0; f < F; ++f)
    for (int b = 0; b < B; ++b) {
      const int64_t bag = (int64_t)f * B + b;
      const int64_t s = offsets[bag], e = offsets[bag + 1];
      const int D = dim[f];
      float* o = out + (int64_t)b * ld_out + col[f];
      for (int d = 0; d < D; ++d) o[d] = 0.f;
      for (int64_t l = s; l < e; ++l) {
        int64_t id = ids[l];
        if (id < 0 || id >= rows[f]) id = 0;
        const float* w = weights + w_off[f] + id * D;
        for (int d = 0; d < D; ++d) o[d] += w[d];
      }
      if (pool[f] == POOL_MEAN && e > s) {
        const float inv = 1.0f / (float)(e - s);
        for (int d = 0; d < D; ++d) o[d] *= inv;
      }
    }
}

/* EXACT fused update.  Thread t owns the keys with key % T == t: it walks all contributions in id order
 * (= the stable order of the oracle), accumulates per unique key in a private open-addressing table, then
 * applies one optimizer update per key.  Deterministic and race-free. */
void orc_fused_update(int32_t optimizer, const float* grad_out, int64_t ld_grad, const int64_t* w_off,
                      const int64_t* rows, const int32_t* dim, const int32_t* col, const int32_t* pool,
                      const int64_t* key_base, const int64_t* ids, const int64_t* offsets, int32_t F, int32_t B,
                      float* weights, float* state, float lr, float eps, float grad_scale, int32_t max_dim) {
  const int64_t nnz = offsets[(int64_t)F * B];
  if (nnz == 0) return;
  /* flatten: key, feature, bag of every id position */
  int64_t* key = (int64_t*)malloc(sizeof(int64_t) * nnz);
  int32_t* bagv = (int32_t*)malloc(sizeof(int32_t) * nnz);
#pragma omp parallel for schedule(static)
  for (int64_t bag = 0; bag < (int64_t)F * B; ++bag) {
    const int f = (int)(bag / B);
    for (int64_t l = offsets[bag]; l < offsets[bag + 1]; ++l) {
      int64_t id = ids[l];
      if (id < 0 || id >= rows[f]) id = 0;
      key[l] = key_base[f] + id;
      bagv[l] = (int32_t)bag;
    }
  }
#pragma omp parallel
  {
    const int T = omp_get_num_threads(), t = omp_get_thread_num();
    int64_t mine = 0;
    for (int64_t l = 0; l < nnz; ++l) mine += ((uint64_t)key[l] % (uint64_t)T) == (uint64_t)t;
    if (mine > 0) {
      int64_t cap = 16;
      while (cap < 2 * mine) cap <<= 1;
      int64_t* hkey = (int64_t*)malloc(sizeof(int64_t) * cap);
      int32_t* hslot = (int32_t*)malloc(sizeof(int32_t) * cap);
      int32_t* hfeat = (int32_t*)malloc(sizeof(int32_t) * mine);
      int64_t* ukey = (int64_t*)malloc(sizeof(int64_t) * mine);
      float* acc = (float*)calloc((size_t)mine * max_dim, sizeof(float));
      for (int64_t i = 0; i < cap; ++i) hkey[i] = -1;
      int64_t n_u = 0;
      for (int64_t l = 0; l < nnz; ++l) {
        const int64_t k = key[l];
        if (((uint64_t)k % (uint64_t)T) != (uint64_t)t) continue;
        int64_t h = (int64_t)(((uint64_t)k * 0x9E3779B97F4A7C15ull) >> 20) & (cap - 1);
        while (hkey[h] != -1 && hkey[h] != k) h = (h + 1) & (cap - 1);
        const int32_t bag = bagv[l];
        const int f = bag / B, b = bag - f * B;
        if (hkey[h] == -1) {
          hkey[h] = k;
          hslot[h] = (int32_t)n_u;
          hfeat[n_u] = f;
          ukey[n_u] = k;
          ++n_u;
        }
        float* a = acc + (int64_t)hslot[h] * max_dim;
        const float* g = grad_out + (int64_t)b * ld_grad + col[f];
        float sc = grad_scale;
        if (pool[f] == POOL_MEAN) sc = grad_scale / (float)(offsets[bag + 1] - offsets[bag]);
        const int D = dim[f];
        for (int d = 0; d < D; ++d) a[d] += g[d] * sc;
      }
      for (int64_t u = 0; u < n_u; ++u) {
        const int f = hfeat[u];
        const int D = dim[f];
        const int64_t row = ukey[u] - key_base[f];
        float* w = weights + w_off[f] + row * D;
        const float* g = acc + u * max_dim;
        if (optimizer == OPT_SGD) {
          for (int d = 0; d < D; ++d) w[d] -= lr * g[d];
        } else if (optimizer == OPT_ADAGRAD) {
          float* s = state + w_off[f] + row * D;
          for (int d = 0; d < D; ++d) {
            const float sn = s[d] + g[d] * g[d];
            s[d] = sn;
            w[d] -= lr * g[d] / (sqrtf(sn) + eps);
          }
        } else {
          float ss = 0.f;
          for (int d = 0; d < D; ++d) ss += g[d] * g[d];
          const float sn = state[ukey[u]] + ss / (float)D;
          state[ukey[u]] = sn;
          const float den = sqrtf(sn) + eps;
          for (int d = 0; d < D; ++d) w[d] -= lr * g[d] / den;
        }
      }
      free(hkey); free(hslot); free(hfeat); free(ukey); free(acc);
    }
  }
  free(key);
  free(bagv);
}

static inline int tri_index(int i, int j, int N) { return i * N - (i * (i + 1)) / 2 + (j - i - 1); }

/* out[b] = [ triu(X X^T, 1) | dense (opt) | sparse (opt) ],  X = [dense ; sparse] */
void orc_dot_interact_fwd(const float* dense, int64_t ld_dense, const float* sparse, int64_t ld_sparse, int64_t B,
                          int32_t Ns, int32_t D, int32_t copy_dense, int32_t copy_sparse, float* out,
                          int64_t ld_out) {
  const int N = Ns + (dense != NULL);
  const int P = N * (N - 1) / 2;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    const float* x[128];
    int n = 0;
    if (dense) x[n++] = dense + b * ld_dense;
    for (int i = 0; i < Ns; ++i) x[n++] = sparse + b * ld_sparse + (int64_t)i * D;
    float* o = out + b * ld_out;
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) {
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc += x[i][d] * x[j][d];
        o[tri_index(i, j, N)] = acc;
      }
    int off = P;
    if (copy_dense && dense) { memcpy(o + off, dense + b * ld_dense, sizeof(float) * D); off += D; }
    if (copy_sparse) memcpy(o + off, sparse + b * ld_sparse, sizeof(float) * (size_t)Ns * D);
  }
}

void orc_dot_interact_bwd(const float* dense, int64_t ld_dense, const float* sparse, int64_t ld_sparse,
                          const float* d_out, int64_t ld_dout, int64_t B, int32_t Ns, int32_t D,
                          int32_t copy_dense, int32_t copy_sparse, float* d_dense, int64_t ld_ddense,
                          float* d_sparse, int64_t ld_dsparse) {
  const int N = Ns + (dense != NULL);
  const int P = N * (N - 1) / 2;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    const float* x[128];
    float* dx[128];
    int n = 0;
    if (dense) { x[n] = dense + b * ld_dense; dx[n++] = d_dense + b * ld_ddense; }
    for (int i = 0; i < Ns; ++i) { x[n] = sparse + b * ld_sparse + (int64_t)i * D; dx[n++] = d_sparse + b * ld_dsparse + (int64_t)i * D; }
    const float* go = d_out + b * ld_dout;
    for (int i = 0; i < N; ++i)
      for (int d = 0; d < D; ++d) dx[i][d] = 0.f;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        if (i == j) continue;
        const float g = go[i < j ? tri_index(i, j, N) : tri_index(j, i, N)];
        for (int d = 0; d < D; ++d) dx[i][d] += g * x[j][d];
      }
    int off = P;
    if (dense && copy_dense) { for (int d = 0; d < D; ++d) dx[0][d] += go[off + d]; off += D; }
    if (copy_sparse) {
      float* ds = d_sparse + b * ld_dsparse;
      for (int k = 0; k < Ns * D; ++k) ds[k] += go[off + k];
    }
  }
}

void orc_fm_fwd(const float* x, int64_t ld_x, int64_t B, int32_t N, int32_t D, float* y, int64_t ld_y) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int d = 0; d < D; ++d) {
      float s = 0.f, q = 0.f;
      for (int n = 0; n < N; ++n) { const float v = x[b * ld_x + (int64_t)n * D + d]; s += v; q += v * v; }
      y[b * ld_y + d] = 0.5f * (s * s - q);
    }
}

void orc_fm_bwd(const float* x, int64_t ld_x, const float* dy, int64_t ld_dy, int64_t B, int32_t N, int32_t D,
                float* dx, int64_t ld_dx) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int d = 0; d < D; ++d) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += x[b * ld_x + (int64_t)n * D + d];
      const float g = dy[b * ld_dy + d];
      for (int n = 0; n < N; ++n) dx[b * ld_dx + (int64_t)n * D + d] = g * (s - x[b * ld_x + (int64_t)n * D + d]);
    }
}
