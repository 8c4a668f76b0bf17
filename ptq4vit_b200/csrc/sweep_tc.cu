// The hot kernel: candidate sweep on tcgen05 tensor cores (sm_100a).
//
// Replaces, for one search step, the reference's loop
//   for c in candidates: out = F.linear(x_sim, w_sim_c) ; sim = -(g*(y-out))**2 ; mean/sum
// (quant_layers/linear.py:466-488, :507-526; quant_layers/matmul.py:500-514, :541-555)
// without ever writing a candidate output to HBM.
//
// One persistent CTA per SM.  Work = (output tile 128x128) x (candidate range),
// split stream-K style over the CTAs.  Per tile fragment:
//   1. the 256 epilogue threads load r = y - bias and g = grad * 2^k for their
//      (row, 64 columns) into REGISTERS -- they stay there for all candidates;
//   2. "fixed" segments (everything that does not change with the candidate) are
//      multiplied on the tensor cores and subtracted: r -= scale * acc;
//   3. per candidate only the segment(s) touched by the candidate step size are
//      multiplied (TMA bulk copy -> smem ring -> tcgen05.mma -> TMEM), and the
//      epilogue forms (g * (r - scale_c * acc))^2 straight from TMEM, reduces it over
//      the 32 rows of the warp with shuffles and writes one partial per 16 columns.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 4..11 = epilogue
// (setmaxnreg moves the register budget of warpgroup 0 to the epilogue warpgroups).
#include "common.cuh"
#include <cstdio>

namespace {

constexpr int kStages = 4;
constexpr int kStageBytes = P4V_TILE * P4V_JOB_KB;       // 16 KB per operand
constexpr int kAccSlots = 3;
constexpr int kAccCols = 128;
constexpr int kTmemCols = 512;
constexpr int kEpiThreads = 256;
constexpr int kThreads = 128 + kEpiThreads;   // warpgroup 0: producer, MMA, 2 idle warps; warpgroups 1-2: epilogue

struct SmemLayout {
  alignas(128) uint8_t stageR[kStages][kStageBytes];
  alignas(128) uint8_t stageC[kStages][kStageBytes];
  alignas(16) P4VJob jobs[P4V_MAX_JOBS];
  float fixs[P4V_MAX_GROUPS][P4V_TILE_CG];
  float candA[P4V_MAX_CAND][P4V_TILE_CG];
  float candB[P4V_MAX_GROUPS][P4V_TILE_CG];
  alignas(8) unsigned long long full[kStages];
  unsigned long long empty[kStages];
  unsigned long long acc_full[kAccSlots];
  unsigned long long acc_empty[kAccSlots];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch (cudaErrorLaunchFailure), never as a hung GPU.
[[noreturn]] __device__ __noinline__ void mbar_timeout(uint32_t addr, uint32_t parity) {
  printf("ptq4vit_b200 sweep: mbarrier wait timed out (block %d thread %d smem 0x%x parity %u)\n",
         (int)blockIdx.x, (int)threadIdx.x, addr, parity);
  __trap();
  while (true) {}
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(addr, parity))
    if (clock64() - t0 > 4000000000ll) mbar_timeout(addr, parity);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(void* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B; LBO = stride between the two
// 16-byte chunks of one K-step, SBO = stride between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  constexpr uint64_t lbo = (P4V_TILE * 16) >> 4;   // 2048 B
  constexpr uint64_t sbo = 128 >> 4;               // 128 B
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
template <bool kInt8>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate) {
  // instruction descriptor: c_format (S32=2 | F32=1) @4, a/b format (S8=1 | BF16=1) @7/@10,
  // K-major both, N>>3 @17, M>>4 @24
  constexpr uint32_t idesc = ((kInt8 ? 2u : 1u) << 4) | (1u << 7) | (1u << 10) |
                             ((uint32_t)(P4V_TILE >> 3) << 17) | ((uint32_t)(P4V_TILE >> 4) << 24);
  if constexpr (kInt8) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
        "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
        "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
        "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct Frag { int tile, p, tm, tn, c0, c1; };

__device__ __forceinline__ bool next_frag(const SweepParams& P, long long& u, long long u_end, Frag& f) {
  if (u >= u_end) return false;
  f.tile = (int)(u / P.n_cand);
  f.c0 = (int)(u % P.n_cand);
  long long rem = u_end - u;
  f.c1 = (int)((rem < (long long)(P.n_cand - f.c0)) ? f.c0 + rem : P.n_cand);
  int per_p = P.tiles_m * P.tiles_n;
  f.p = f.tile / per_p;
  int t = f.tile % per_p;
  if (P.order == 0) { f.tm = t % P.tiles_m; f.tn = t / P.tiles_m; }
  else              { f.tn = t % P.tiles_n; f.tm = t / P.tiles_n; }
  u += f.c1 - f.c0;
  return true;
}

template <bool kInt8>
__device__ __forceinline__ float acc_to_float(uint32_t a) {
  if constexpr (kInt8) return __int2float_rn((int)a);
  else return __uint_as_float(a);
}

// Reduce 4 per-lane partial sums over the 32 lanes (= 32 rows).  After the call lane
// 8*k (k=0..3) holds the total of value k.  Fixed order => deterministic.
__device__ __forceinline__ float reduce4_over_rows(float v0, float v1, float v2, float v3, int lane) {
  const unsigned full = 0xffffffffu;
  bool hi16 = lane & 16;
  float s0 = hi16 ? v0 : v2, s1 = hi16 ? v1 : v3;      // what this lane sends away
  float k0 = hi16 ? v2 : v0, k1 = hi16 ? v3 : v1;      // what it keeps
  k0 += __shfl_xor_sync(full, s0, 16);
  k1 += __shfl_xor_sync(full, s1, 16);
  bool hi8 = lane & 8;
  float s = hi8 ? k0 : k1, k = hi8 ? k1 : k0;
  k += __shfl_xor_sync(full, s, 8);
  k += __shfl_xor_sync(full, k, 4);
  k += __shfl_xor_sync(full, k, 2);
  k += __shfl_xor_sync(full, k, 1);
  return k;   // lanes with (bit4,bit3) = (a,b) hold value 2a+b
}

template <bool kInt8, bool kSingle>
__global__ void __launch_bounds__(kThreads, 1) sweep_tc_kernel(const __grid_constant__ SweepParams P) {
  extern __shared__ uint8_t smem_raw[];
  SmemLayout& S = *reinterpret_cast<SmemLayout*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup ----
  const int n_jobs = P.n_fixed_jobs + P.n_cand_jobs;
  for (int i = threadIdx.x; i < n_jobs; i += kThreads) S.jobs[i] = P.jobs[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
    for (int i = 0; i < kAccSlots; ++i) { mbar_init(&S.acc_full[i], 1); mbar_init(&S.acc_empty[i], kEpiThreads); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  const long long units = (long long)P.P * P.tiles_m * P.tiles_n * P.n_cand;
  long long u = units * blockIdx.x / gridDim.x;
  const long long u_end = units * (blockIdx.x + 1) / gridDim.x;
  Frag f;

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      auto issue = [&](const P4VJob& j, const Frag& fr, int c) {
        mbar_wait(&S.empty[stage], phase ^ 1);
        const uint32_t bytes = (uint32_t)j.kb * P4V_TILE;
        const size_t rt = (size_t)(fr.p * P.tiles_m + fr.tm), ct = (size_t)(fr.p * P.tiles_n + fr.tn);
        const uint8_t* r = ((j.flags & P4V_JOB_RCAND) ? P.R_cand + (size_t)c * P.R_cand_stride + rt * P.R_cand_tile_bytes
                                                      : P.R_cur + rt * P.R_tile_bytes) + j.r_off;
        const uint8_t* cc = ((j.flags & P4V_JOB_CCAND) ? P.C_cand + (size_t)c * P.C_cand_stride + ct * P.C_cand_tile_bytes
                                                       : P.C_cur + ct * P.C_tile_bytes) + j.c_off;
        mbar_expect_tx(&S.full[stage], 2 * bytes);
        bulk_g2s(S.stageR[stage], r, bytes, &S.full[stage]);
        bulk_g2s(S.stageC[stage], cc, bytes, &S.full[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      };
      while (next_frag(P, u, u_end, f)) {
        for (int j = 0; j < P.n_fixed_jobs; ++j) issue(S.jobs[j], f, 0);
        for (int c = f.c0; c < f.c1; ++c)
          for (int j = 0; j < P.n_cand_jobs; ++j) issue(S.jobs[P.n_fixed_jobs + j], f, c);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, slot = 0, slot_phase = 0;
      auto run = [&](const P4VJob& j) {
        if (j.flags & P4V_JOB_FIRST) {
          mbar_wait(&S.acc_empty[slot], slot_phase ^ 1);
          tc_fence_after();
        }
        mbar_wait(&S.full[stage], phase);
        tc_fence_after();
        const uint32_t ra = smem_u32(S.stageR[stage]), ca = smem_u32(S.stageC[stage]);
        const uint32_t d = tmem + kAccCols + slot * kAccCols;
        const int ksteps = j.kb >> 5;
        for (int k = 0; k < ksteps; ++k) {
          uint32_t acc = ((j.flags & P4V_JOB_FIRST) && k == 0) ? 0u : 1u;
          umma<kInt8>(d, make_desc(ra + k * 2 * P4V_TILE * 16), make_desc(ca + k * 2 * P4V_TILE * 16), acc);
        }
        tc_commit(&S.empty[stage]);
        if (j.flags & P4V_JOB_LAST) {
          tc_commit(&S.acc_full[slot]);
          if (++slot == kAccSlots) { slot = 0; slot_phase ^= 1; }
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      };
      while (next_frag(P, u, u_end, f)) {
        for (int j = 0; j < P.n_fixed_jobs; ++j) run(S.jobs[j]);
        for (int c = f.c0; c < f.c1; ++c)
          for (int j = 0; j < P.n_cand_jobs; ++j) run(S.jobs[P.n_fixed_jobs + j]);
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ======================= epilogue (8 warps) =======================
    const int ew = warp - 4;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int hf = ew >> 2;                  // column half
    const int et = threadIdx.x - 128;        // 0..255
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float gs = P.out ? 1.f : *P.gscale;
    uint32_t slot = 0, slot_phase = 0;
    float r[64], g[64];

    while (next_frag(P, u, u_end, f)) {
      // -- scale tables for this tile's 8 column groups --
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));   // previous fragment done with the tables
      {
        const int sg0 = (P.sg_mode == P4V_SG_COLUMN) ? f.tn * P4V_TILE_CG : (f.p % P.nsg);
        const int sgs = (P.sg_mode == P4V_SG_COLUMN) ? 1 : 0;
        for (int i = et; i < P.n_fixed_groups * P4V_TILE_CG; i += kEpiThreads)
          S.fixs[i >> 3][i & 7] = P.fix_scale[(size_t)(i >> 3) * P.nsg + sg0 + (i & 7) * sgs];
        for (int i = et; i < P.n_cand_groups * P4V_TILE_CG; i += kEpiThreads)
          S.candB[i >> 3][i & 7] = P.candB[(size_t)(i >> 3) * P.nsg + sg0 + (i & 7) * sgs];
        for (int i = et + f.c0 * P4V_TILE_CG; i < f.c1 * P4V_TILE_CG; i += kEpiThreads)
          S.candA[i >> 3][i & 7] = P.candA[(size_t)(i >> 3) * P.nsg + sg0 + (i & 7) * sgs];
      }
      // -- residual target and gradient tile into registers --
      {
        const int gm = f.tm * P4V_TILE + quarter * 32 + lane;
        const int col0 = f.tn * P4V_TILE + hf * 64;
        const float* yrow = P.Y + (size_t)f.p * P.prob_stride + (size_t)gm * P.ld;
        const float* grow = P.Gr + (size_t)f.p * P.prob_stride + (size_t)gm * P.ld;
        const bool row_ok = gm < P.M;
        if (P.out != nullptr) {           // quant_forward: r starts at -bias, output = -r
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const int col = col0 + j;
            r[j] = (P.bias && col < P.N) ? -P.bias[col] : 0.f;
            g[j] = 0.f;
          }
        } else if (row_ok && (P.ld & 3) == 0 && col0 + 64 <= P.N) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            float4 yv = *reinterpret_cast<const float4*>(yrow + col0 + j);
            float4 gv = *reinterpret_cast<const float4*>(grow + col0 + j);
            float4 bv = P.bias ? *reinterpret_cast<const float4*>(P.bias + col0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            r[j] = yv.x - bv.x; r[j + 1] = yv.y - bv.y; r[j + 2] = yv.z - bv.z; r[j + 3] = yv.w - bv.w;
            g[j] = gv.x * gs; g[j + 1] = gv.y * gs; g[j + 2] = gv.z * gs; g[j + 3] = gv.w * gs;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const int col = col0 + j;
            const bool ok = row_ok && col < P.N;
            r[j] = ok ? (yrow[col] - (P.bias ? P.bias[col] : 0.f)) : 0.f;
            g[j] = ok ? grow[col] * gs : 0.f;
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads));   // tables visible

      // -- fixed segments: r -= scale * acc --
      for (int gi = 0; gi < P.n_fixed_groups; ++gi) {
        mbar_wait(&S.acc_full[slot], slot_phase);
        tc_fence_after();
        const uint32_t taddr = tmem + lane_addr + kAccCols + slot * kAccCols + hf * 64;
        const float4 sc = *reinterpret_cast<const float4*>(&S.fixs[gi][hf * 4]);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t a[32];
          tmem_ld32(taddr + ch * 32, a);
          tmem_wait_ld();
          const float s_lo = ch ? sc.z : sc.x, s_hi = ch ? sc.w : sc.y;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            r[ch * 32 + j] = fmaf(-(j < 16 ? s_lo : s_hi), acc_to_float<kInt8>(a[j]), r[ch * 32 + j]);
        }
        tc_fence_before();
        mbar_arrive(&S.acc_empty[slot]);
        if (++slot == kAccSlots) { slot = 0; slot_phase ^= 1; }
      }
      if (P.out != nullptr) {
        const int gm = f.tm * P4V_TILE + quarter * 32 + lane;
        const int col0 = f.tn * P4V_TILE + hf * 64;
        if (gm < P.M) {
          float* orow = P.out + (size_t)f.p * P.prob_stride + (size_t)gm * P.ld;
#pragma unroll
          for (int j = 0; j < 64; ++j) if (col0 + j < P.N) orow[col0 + j] = -r[j];
        }
        continue;
      }
      if constexpr (!kSingle) {   // park the residual target in TMEM columns [0,128)
        tmem_st32(tmem + lane_addr + hf * 64, r);
        tmem_st32(tmem + lane_addr + hf * 64 + 32, r + 32);
        tmem_wait_st();
      }

      // -- candidates --
      float* part_base = P.partial + ((size_t)f.tile * P.n_cand) * 32 + quarter * 8 + hf * 4;
      for (int c = f.c0; c < f.c1; ++c) {
        const float4 ca = *reinterpret_cast<const float4*>(&S.candA[c][hf * 4]);
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if constexpr (!kSingle) {
          uint32_t t[32];
          tmem_ld32(tmem + lane_addr + hf * 64, t);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(t[j]);
          tmem_ld32(tmem + lane_addr + hf * 64 + 32, t);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[32 + j] = __uint_as_float(t[j]);
        }
        for (int gi = 0; gi < P.n_cand_groups; ++gi) {
          mbar_wait(&S.acc_full[slot], slot_phase);
          tc_fence_after();
          const uint32_t taddr = tmem + lane_addr + kAccCols + slot * kAccCols + hf * 64;
          const float4 cb = *reinterpret_cast<const float4*>(&S.candB[gi][hf * 4]);
          const bool noA = (P.cand_noA_mask >> gi) & 1ull;
          const float4 sc = noA ? cb : make_float4(ca.x * cb.x, ca.y * cb.y, ca.z * cb.z, ca.w * cb.w);
          const bool last = kSingle || (gi == P.n_cand_groups - 1);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t a[32];
            tmem_ld32(taddr + ch * 32, a);
            tmem_wait_ld();
            const float s_lo = ch ? sc.z : sc.x, s_hi = ch ? sc.w : sc.y;
            if (last) {
              float q_lo = 0.f, q_hi = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float d = fmaf(-(j < 16 ? s_lo : s_hi), acc_to_float<kInt8>(a[j]), r[ch * 32 + j]);
                const float w = g[ch * 32 + j] * d;
                if (j < 16) q_lo = fmaf(w, w, q_lo); else q_hi = fmaf(w, w, q_hi);
              }
              if (ch == 0) { p0 = q_lo; p1 = q_hi; } else { p2 = q_lo; p3 = q_hi; }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                r[ch * 32 + j] = fmaf(-(j < 16 ? s_lo : s_hi), acc_to_float<kInt8>(a[j]), r[ch * 32 + j]);
            }
          }
          tc_fence_before();
          mbar_arrive(&S.acc_empty[slot]);
          if (++slot == kAccSlots) { slot = 0; slot_phase ^= 1; }
        }
        const float tot = reduce4_over_rows(p0, p1, p2, p3, lane);
        if ((lane & 7) == 0) part_base[(size_t)c * 32 + (lane >> 3)] = tot;
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
  }
}

}  // namespace

int p4v_launch_sweep_tc(const SweepParams& p, int num_sms, cudaStream_t st) {
  P4V_REQUIRE(p.n_fixed_jobs + p.n_cand_jobs <= P4V_MAX_JOBS, "sweep: too many jobs (%d)", p.n_fixed_jobs + p.n_cand_jobs);
  P4V_REQUIRE(p.n_fixed_groups <= P4V_MAX_GROUPS && p.n_cand_groups <= P4V_MAX_GROUPS, "sweep: too many segment groups");
  P4V_REQUIRE(p.n_cand <= P4V_MAX_CAND && p.n_cand >= 1, "sweep: bad candidate count");
  P4V_REQUIRE(p.out != nullptr ? (p.n_cand == 1 && p.n_cand_jobs == 0) : p.n_cand_groups >= 1, "sweep: bad mode");
  const long long units = (long long)p.P * p.tiles_m * p.tiles_n * p.n_cand;
  int grid = (int)(units < num_sms ? units : num_sms);
  if (grid < 1) return 0;
  const size_t smem = sizeof(SmemLayout) + 128;
  const bool single = p.n_cand_groups == 1;
#define P4V_LAUNCH(I8, SG)                                                                             \
  do {                                                                                                 \
    P4V_CUDA_OK(cudaFuncSetAttribute(sweep_tc_kernel<I8, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    sweep_tc_kernel<I8, SG><<<grid, kThreads, smem, st>>>(p);                                          \
  } while (0)
  if (p.is_int8) { if (single) P4V_LAUNCH(true, true); else P4V_LAUNCH(true, false); }
  else           { if (single) P4V_LAUNCH(false, true); else P4V_LAUNCH(false, false); }
#undef P4V_LAUNCH
  P4V_CUDA_OK(cudaGetLastError());
  return 0;
}
