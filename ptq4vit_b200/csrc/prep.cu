#include "prep.cuh"
#include <math.h>

void p4v_count_launch();

namespace {

// order-preserving float <-> int key (so that atomicMax on ints is max on floats)
__device__ __forceinline__ int f2key(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key2f(int k) {
  return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}
constexpr int kKeyMin = (int)0x80000000;

__global__ void keys_reset_kernel(int* keys, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = kKeyMin;
}

// grid (row_split, n_col_blocks, n_row_blocks)
__global__ void block_max_kernel(const float* __restrict__ src, long long ld, int rows, int row_block,
                                 int col_block, int use_abs, int* keys) {
  const int rb = blockIdx.z, cb = blockIdx.y;
  const int r_begin = rb * row_block, r_end = min(rows, r_begin + row_block);
  const int rows_here = max(0, r_end - r_begin);
  const int per = (rows_here + gridDim.x - 1) / gridDim.x;
  const int r0 = r_begin + blockIdx.x * per, r1 = min(r_end, r0 + per);
  float m = -INFINITY;
  for (int r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    const float* row = src + (size_t)r * ld + (size_t)cb * col_block;
    for (int c = threadIdx.x; c < col_block; c += blockDim.x) {
      float v = row[c];
      m = fmaxf(m, use_abs ? fabsf(v) : v);
    }
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) {
    const int nw = (blockDim.x * blockDim.y + 31) / 32;
    for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]);
    if (m > -INFINITY) atomicMax(&keys[rb * gridDim.y + cb], f2key(m));
  }
}

// absmax over all problems p with p % n_groups == g ; grid (split, n_groups)
__global__ void group_absmax_kernel(const float* __restrict__ src, long long prob_elems, int P, int n_groups, int* keys) {
  const int g = blockIdx.y;
  float m = 0.f;
  for (int p = g; p < P; p += n_groups) {
    const float* base = src + (size_t)p * prob_elems;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < prob_elems; i += (long long)gridDim.x * blockDim.x)
      m = fmaxf(m, fabsf(base[i]));
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x + 31) / 32; ++i) m = fmaxf(m, red[i]);
    atomicMax(&keys[g], f2key(m));
  }
}

__global__ void keys_to_delta_kernel(const int* keys, int n, float denom, float* d0, float* d1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = __fdiv_rn(key2f(keys[i]), denom);
    d0[i] = v;
    if (d1) d1[i] = v;
  }
}

__global__ void make_gscale_kernel(const int* key, float* gscale) {
  float m = key2f(key[0]);
  float s = 1.f;
  if (m > 0.f && isfinite(m)) {
    int e;
    frexpf(m, &e);             // m = f * 2^e, f in [0.5,1)
    e = max(-100, min(100, 1 - e));
    s = ldexpf(1.f, e);        // m * s in [1,2)
  }
  gscale[0] = s;
}

// One thread = one (plane, problem, padded row, 16-byte chunk of one segment).
template <bool kInt8>
__global__ void quant_image_kernel(const QuantImageArgs a, int chunks_total) {
  const long long rows_pad = (long long)a.tiles * P4V_TILE;
  const long long per_plane = (long long)a.P * rows_pad * chunks_total;
  const long long total = per_plane * a.n_planes;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    // row fastest so that the 16-byte stores of a warp are contiguous
    const int plane = (int)(idx / per_plane);
    long long rem = idx % per_plane;
    const int row_p = (int)(rem % rows_pad); rem /= rows_pad;
    const int chunk_g = (int)(rem % chunks_total);
    const int p = (int)(rem / chunks_total);
    // find the segment of this chunk
    int s = 0, chunk = chunk_g;
    constexpr int epc = kInt8 ? 16 : 8;          // elements per 16-byte chunk
    while (true) {
      const int nch = ((a.segs[s].klen + (kInt8 ? 31 : 15)) / (kInt8 ? 32 : 16)) * 2;   // chunks of this segment (padded to 32 B)
      if (chunk < nch) break;
      chunk -= nch; ++s;
    }
    const P4VSeg sg = a.segs[s];
    const int tile = row_p / P4V_TILE, r = row_p % P4V_TILE;
    uint8_t* dst = a.dst + (size_t)plane * a.plane_stride + ((size_t)p * a.tiles + tile) * a.tile_bytes + sg.dst_off +
                   ((size_t)chunk * P4V_TILE + r) * 16;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (row_p < a.rows) {
      float delta = 1.f;
      const float split = sg.sos_part ? (a.factors ? a.factors[plane] : a.split[0]) : 0.f;
      if (sg.sos_part || sg.split3) { /* step size handled below */ }
      else if (sg.fixed_delta > 0.f) delta = sg.fixed_delta;
      else {
        const int rb = a.rows_per_block > 0 ? row_p / a.rows_per_block : (p % a.d_mod);
        delta = a.delta[(size_t)rb * a.d_stride + sg.didx];
        if (a.factors) delta = a.factors[plane] * delta;       // fl(f_c * delta0), as the reference's candidate table
      }
      const float* base = a.src + (size_t)p * a.prob_stride;
#pragma unroll
      for (int e = 0; e < epc; ++e) {
        const int kk = chunk * epc + e;
        float q = 0.f;
        if (kk < sg.klen) {
          const int k = sg.k0 + kk;
          const float v = a.src_transposed ? base[(size_t)k * a.ld + row_p] : base[(size_t)row_p * a.ld + k];
          if (sg.split3) {
            const float b1 = __bfloat162float(__float2bfloat16_rn(v));
            const float b2 = __bfloat162float(__float2bfloat16_rn(v - b1));
            q = sg.split3 == 1 ? b1 : (sg.split3 == 2 ? b2 : __bfloat162float(__float2bfloat16_rn((v - b1) - b2)));
          } else if (sg.sos_part == 1) {
            q = fminf(fmaxf(rintf(fminf(fmaxf(v, split), 1.f) * sg.qm1), 0.f), sg.qm1);
          } else if (sg.sos_part == 2) {
            q = fminf(fmaxf(rintf(__fdiv_rn(fminf(fmaxf(v, 0.f), split), __fdiv_rn(split, sg.qm1))), 0.f), sg.qm1);
          } else {
            q = fminf(fmaxf(rintf(__fdiv_rn(v, delta)), sg.lo), sg.hi);
          }
          if (!(q == q)) q = 0.f;   // NaN (0/0) cannot be represented in the integer operand
        }
        if constexpr (kInt8) {
          const int qi = (int)q;
          w[e >> 2] |= (uint32_t)(qi & 0xff) << ((e & 3) * 8);
        } else {
          const uint32_t hb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(q));
          w[e >> 1] |= hb << ((e & 1) * 16);
        }
      }
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void step_tables_kernel(const StepTablesArgs a) {
  const int total_fix = a.n_fixed_groups * a.nsg;
  const int total_cb = a.n_cand_groups * a.nsg;
  const int total_ca = a.n_cand * a.nsg;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_fix + total_cb + total_ca; i += gridDim.x * blockDim.x) {
    if (i < total_fix) {
      const int g = i / a.nsg, sg = i % a.nsg;
      const GroupMeta m = a.fixed_meta[g];
      if (a.kind >= 2) {
        a.fix_scale[i] = a.dW[sg] * (a.kind == 2 ? a.dX[sg] : a.dX[m.a]);
      } else {
        const int v = min((sg * P4V_CG) / a.crb_rows, a.n_V - 1);
        a.fix_scale[i] = a.dW[v * a.n_H + m.h] * (m.neg ? a.d_neg : a.dX[m.a]);
      }
    } else if (i < total_fix + total_cb) {
      const int j = i - total_fix;
      const int g = j / a.nsg, sg = j % a.nsg;
      const GroupMeta m = a.cand_meta[g];
      float val;
      if (a.kind == 0)      val = m.neg ? a.d_neg : a.dX[m.a];
      else if (a.kind == 1) val = a.dW[min((sg * P4V_CG) / a.crb_rows, a.n_V - 1) * a.n_H + m.h];
      else if (a.kind == 2) val = a.dX[sg];                         // head-wise other operand
      else                  val = m.neg ? a.d_neg : a.dX[m.a];      // uniform other-operand scale per group
      a.candB[j] = val;
    } else {
      const int j = i - total_fix - total_cb;
      const int c = j / a.nsg, sg = j % a.nsg;
      float base;
      if (a.kind == 0)      base = a.dW0[min((sg * P4V_CG) / a.crb_rows, a.n_V - 1) * a.n_H + a.target];
      else if (a.kind == 1) base = a.dX0[a.target];
      else                  base = a.dW0[sg];
      a.candA[j] = a.factors[c] * base;
    }
  }
}

// grid (n_cand, n_groups), 128 threads.  Fixed item order + fixed tree => deterministic.
__global__ void reduce_scores_kernel(const ReduceArgs a) {
  const int c = blockIdx.x, g = blockIdx.y;
  const int per_p = a.tiles_m * a.tiles_n;
  double acc = 0.0;
  if (a.mode == P4V_SG_COLUMN) {
    const int ncg = a.tiles_n * P4V_TILE_CG;
    const int cg0 = a.n_groups == 1 ? 0 : g * a.cg_per_group;
    const int cg1 = a.n_groups == 1 ? ncg : min(ncg, cg0 + a.cg_per_group);
    const int ncgs = cg1 - cg0;
    const long long items = (long long)a.P * a.tiles_m * 4 * ncgs;
    for (long long i = threadIdx.x; i < items; i += blockDim.x) {
      const int cgl = (int)(i % ncgs); long long r = i / ncgs;
      const int q = (int)(r % 4); r /= 4;
      const int tm = (int)(r % a.tiles_m); const int p = (int)(r / a.tiles_m);
      const int cg = cg0 + cgl, tn = cg >> 3, i8 = cg & 7;
      const int t = a.order == 0 ? tn * a.tiles_m + tm : tm * a.tiles_n + tn;
      acc += (double)a.partial[(((size_t)p * per_p + t) * a.n_cand + c) * 32 + q * 8 + i8];
    }
  } else {
    const int np = (a.P - g + a.n_groups - 1) / a.n_groups;     // problems p = g + k*n_groups
    const long long items = (long long)np * per_p * 32;
    for (long long i = threadIdx.x; i < items; i += blockDim.x) {
      const int e = (int)(i % 32); long long r = i / 32;
      const int t = (int)(r % per_p); const int k = (int)(r / per_p);
      const int p = g + k * a.n_groups;
      acc += (double)a.partial[(((size_t)p * per_p + t) * a.n_cand + c) * 32 + e];
    }
  }
  __shared__ double red[128];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double gs = (double)a.gscale[0];
    a.scores[(size_t)c * a.n_groups + g] = -red[0] * a.inv_count / (gs * gs);
  }
}

// Every block recomputes the (tiny) argmax table; block 0 publishes the new step sizes;
// all blocks copy the winning candidate's image slabs into the current image.
__global__ void finish_step_kernel(const FinishArgs a, int chunks_total) {
  extern __shared__ int s_best[];
  for (int g = threadIdx.x; g < a.n_groups; g += blockDim.x) {
    int bi = 0; double bv = a.scores[g];
    for (int c = 1; c < a.n_cand; ++c) {
      const double v = a.scores[(size_t)c * a.n_groups + g];
      if (v > bv || (v != v && bv == bv)) { bv = v; bi = c; }     // first maximum; NaN wins like torch.argmax
    }
    s_best[g] = bi;
    if (blockIdx.x == 0) {
      const size_t di = (size_t)g * a.d_stride + a.d_col;
      a.d[di] = a.factors[bi] * a.d0[di];
      if (a.best) a.best[g] = bi;
    }
  }
  if (blockIdx.x == 0 && a.score_log)
    for (int i = threadIdx.x; i < a.n_cand * a.n_groups; i += blockDim.x) a.score_log[i] = (float)a.scores[i];
  __syncthreads();
  if (a.nseg == 0) return;
  const long long rows_pad = (long long)a.tiles * P4V_TILE;
  const long long total = (long long)a.P * rows_pad * chunks_total;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int row_p = (int)(idx % rows_pad); long long rem = idx / rows_pad;
    int chunk = (int)(rem % chunks_total); const int p = (int)(rem / chunks_total);
    int s = 0;
    while (chunk >= (a.segs[s].kb >> 4)) { chunk -= a.segs[s].kb >> 4; ++s; }
    int g = 0;
    if (a.problem_groups) g = p % a.n_groups;
    else if (a.rows_per_group > 0) g = min(row_p / a.rows_per_group, a.n_groups - 1);
    const int tile = row_p / P4V_TILE, r = row_p % P4V_TILE;
    const size_t in_tile = ((size_t)chunk * P4V_TILE + r) * 16;
    const uint4 v = *reinterpret_cast<const uint4*>(a.cand + (size_t)s_best[g] * a.cand_plane_stride +
                                                   ((size_t)p * a.tiles + tile) * a.cand_tile_bytes + a.segs[s].src_off + in_tile);
    *reinterpret_cast<uint4*>(a.cur + ((size_t)p * a.tiles + tile) * a.cur_tile_bytes + a.segs[s].dst_off + in_tile) = v;
  }
}

int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

int p4v_keys_reset(int* keys, int n, cudaStream_t st) {
  keys_reset_kernel<<<p4v_cdiv(n, 128), 128, 0, st>>>(keys, n); p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}

int p4v_block_max(const float* src, long long ld, int rows, int row_block, int n_row_blocks, int col_block,
                  int n_col_blocks, int use_abs, int* keys, cudaStream_t st) {
  long long elems = (long long)row_block * col_block;
  int split = (int)((elems + (1 << 16) - 1) >> 16);
  if (split < 1) split = 1;
  if (split > 256) split = 256;
  if (split > row_block) split = row_block;
  dim3 grid(split, n_col_blocks, n_row_blocks), block(32, 8);
  block_max_kernel<<<grid, block, 0, st>>>(src, ld, rows, row_block, col_block, use_abs, keys); p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}

int p4v_group_absmax(const float* src, long long prob_elems, int P, int n_groups, int* keys, cudaStream_t st) {
  long long per_group = prob_elems * ((P + n_groups - 1) / n_groups);
  int split = (int)((per_group + (1 << 16) - 1) >> 16);
  if (split < 1) split = 1;
  if (split > 128) split = 128;
  dim3 grid(split, n_groups);
  group_absmax_kernel<<<grid, 256, 0, st>>>(src, prob_elems, P, n_groups, keys); p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}

int p4v_keys_to_delta(const int* keys, int n, float denom, float* d0, float* d1, cudaStream_t st) {
  keys_to_delta_kernel<<<p4v_cdiv(n, 128), 128, 0, st>>>(keys, n, denom, d0, d1); p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}

int p4v_make_gscale(const int* key, float* gscale, cudaStream_t st) {
  make_gscale_kernel<<<1, 1, 0, st>>>(key, gscale); p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}

int p4v_quant_image(const QuantImageArgs& a, cudaStream_t st) {
  // chunk count per row over all segments (host copy of the seg table is not available here:
  // the caller passes tile_bytes = 128 * padded bytes, and every segment is padded to 32 B)
  const int chunks_total = (int)(a.tile_bytes / P4V_TILE / 16);
  const long long total = (long long)a.n_planes * a.P * a.tiles * P4V_TILE * chunks_total;
  if (total == 0) return 0;
  const int grid = grid_for(total, 256, 148 * 32);
  if (a.is_int8) quant_image_kernel<true><<<grid, 256, 0, st>>>(a, chunks_total);
  else quant_image_kernel<false><<<grid, 256, 0, st>>>(a, chunks_total);
  p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}

int p4v_step_tables(const StepTablesArgs& a, cudaStream_t st) {
  const int total = (a.n_fixed_groups + a.n_cand_groups + a.n_cand) * a.nsg;
  if (total == 0) return 0;
  step_tables_kernel<<<grid_for(total, 256, 64), 256, 0, st>>>(a); p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}

int p4v_reduce_scores(const ReduceArgs& a, cudaStream_t st) {
  dim3 grid(a.n_cand, a.n_groups);
  reduce_scores_kernel<<<grid, 128, 0, st>>>(a); p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}

int p4v_finish_step(const FinishArgs& a, cudaStream_t st) {
  const long long total = a.nseg > 0 ? (long long)a.P * a.tiles * P4V_TILE * a.commit_chunks : 0;
  const int grid = grid_for(total, 256, 148 * 8);
  finish_step_kernel<<<grid, 256, (size_t)a.n_groups * sizeof(int), st>>>(a, a.commit_chunks); p4v_count_launch();
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}
