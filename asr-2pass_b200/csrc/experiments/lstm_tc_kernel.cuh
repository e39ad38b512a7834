// EXPERIMENT, not compiled into libb200pf.so: the tcgen05 variant of the LSTM step that round 1 measured slower than the
// mma.sync kernel in lstm.cu (6.5 vs 5.3 us per step).  Kept as a record of what was tried; to build it,
// paste it back into lstm.cu's anonymous namespace and dispatch to it from lstm_launch.
// ------------------------------------------------------------------------------------------------
// tcgen05 variant (experiment, not the default: see lstm_launch).  Same cluster-of-16 decomposition, but the per-step slab product runs on the 5th-gen
// tensor core: D[128 gate rows x 32 sequences] (TMEM, 32 columns) = W slab [128 x 512] x h_{t-1}^T, both operands K-major
// SWIZZLE_128B tiles in shared memory, 32 tcgen05.mma (M128 N32 K16) per step issued by one thread.  Slab row
// R = 32 gate + unit, so warp `gate` (TMEM lane quarter `gate`) reads that gate for 32 units x 32 sequences, adds the
// input projection, applies its non-linearity (tanh for g, sigmoid otherwise) and parks the result in shared memory;
// after one CTA barrier each thread updates 8 (unit, sequence) cells whose state lives in its registers, writes the
// new hidden values into its own 64-byte piece of h_t (K-major: K block rank/2, row = sequence), and the CTA pushes
// that 2 KB piece to its 15 peers with 16-byte st.shared::cluster.  fence.proxy.async + the cluster barrier make the
// generic-proxy writes visible to the next step's MMAs.
// ------------------------------------------------------------------------------------------------
constexpr int kTcThreads = 128;
constexpr int kTcW = 128 * 1024;        // 8 K blocks x [128 rows x 128 B]
constexpr int kTcH = 32 * 1024;         // 8 K blocks x [32 rows x 128 B], per buffer
constexpr int kTcAct = 4 * 32 * 32 * 4; // [gate][seq][unit] fp32
constexpr int kTcSmem = kTcW + 2 * kTcH + kTcAct + 2 * kGroup * 4 + 64 + 1024 /*align*/;


__global__ void __launch_bounds__(kTcThreads, 1)
lstm_tc_kernel(LstmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sH = smem + kTcW;
  float* sAct = reinterpret_cast<float*>(smem + kTcW + 2 * kTcH);
  int* sOff = reinterpret_cast<int*>(smem + kTcW + 2 * kTcH + kTcAct);
  int* sLen = sOff + kGroup;
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sLen + kGroup);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int task = blockIdx.x / kCluster;
  const int dir = task % p.n_dir, group = task / p.n_dir;
  const bool rev = (p.reverse_mask >> dir) & 1;
  const int s0 = group * kGroup;

  if (tid < kGroup) {
    const int s = s0 + tid;
    sOff[tid] = s < p.n_seq ? p.seq_off[s] : 0;
    sLen[tid] = s < p.n_seq ? p.seq_len[s] : 0;
  }
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 32);
  {
    // slab row R = 32 gate + unit  <-  W_hh row gate*512 + 32 rank + unit;  K block kb, 16-byte chunk j of row R at
    // kb*16384 + R*128 + ((j ^ (R & 7)) << 4)
    const __nv_bfloat16* wsrc = p.whh + (size_t)dir * 2048 * 512;
    for (int idx = tid; idx < 128 * 64; idx += kTcThreads) {
      const int R = idx >> 6, c = idx & 63;          // c: 16-byte chunk along K (64 per row)
      const int grow = (R >> 5) * 512 + (int)rank * 32 + (R & 31);
      const uint4 v = ldg128_nc(wsrc + (size_t)grow * 512 + c * 8);
      const int kb = c >> 3, j = c & 7;
      *reinterpret_cast<uint4*>(sW + kb * 16384 + R * 128 + ((j ^ (R & 7)) << 4)) = v;
    }
    for (int idx = tid; idx < 2 * kTcH / 16; idx += kTcThreads) reinterpret_cast<uint4*>(sH)[idx] = make_uint4(0, 0, 0, 0);
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  int steps = 0;
#pragma unroll 1
  for (int i = 0; i < kGroup; ++i) steps = max(steps, sLen[i]);
  cluster_sync_all();

  const int gate = warp, unit = lane;                 // activation phase: this thread's TMEM lane
  const unsigned short* gx = reinterpret_cast<const unsigned short*>(p.gx);
  const int gcol = dir * 2048 + gate * 512 + (int)rank * 32 + unit;
  // cell phase: unit = lane, sequences 8 warp .. 8 warp + 7
  float c_state[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) c_state[x] = 0.f;
  const uint32_t sW_u = smem_u32(sW), sH_u = smem_u32(sH), sAct_u = smem_u32(sAct);
  constexpr uint32_t idesc = umma_idesc_bf16(128, 32);
  // this CTA's piece of h: K block rank/2, chunks (rank & 1) * 4 .. + 3 of each sequence row
  const uint32_t piece_kb = (rank >> 1) * 4096u, piece_j0 = (rank & 1) * 4u;

  // input-projection terms for (gate, unit) x 32 sequences, fetched ONE STEP AHEAD: the scattered 2-byte loads take
  // longer than a whole step's MMAs, so they must never be waited for inside the step they belong to
  unsigned short gxv[32];
  auto load_gx = [&](int kk, unsigned short (&dst)[32]) {
#pragma unroll
    for (int s = 0; s < 32; ++s) {
      const int len = sLen[s];
      const int t = rev ? len - 1 - kk : kk;
      dst[s] = (kk < len && !(p.dbg & 1)) ? __ldg(gx + (size_t)(sOff[s] + t) * p.ld_gx + gcol) : (unsigned short)0;
    }
  };
  load_gx(0, gxv);
  for (int k = 0; k < steps; ++k) {
    const uint32_t hb = sH_u + (uint32_t)(k & 1) * kTcH;
    const uint32_t hn = sH_u + (uint32_t)((k + 1) & 1) * kTcH;
    if (tid == 0 && !(p.dbg & 16)) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        const uint64_t da = umma_desc_sw128(sW_u + kb * 16384);
        const uint64_t db = umma_desc_sw128(hb + kb * 4096);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) umma_bf16(tmem_d, da + 2 * k16, db + 2 * k16, idesc, (kb | k16) != 0 ? 1u : 0u);
      }
      umma_commit(mma_bar);
    }
    unsigned short gxn[32];
    load_gx(k + 1, gxn);
    if (!(p.dbg & 16)) mbar_wait(mma_bar, (uint32_t)(k & 1));
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    tc_fence_before();
    // tanh(x) = 2 sigmoid(2x) - 1: one code path for all four gates
    const float ka = gate == 2 ? 2.f : 1.f, kb2 = gate == 2 ? -1.f : 0.f;
#pragma unroll
    for (int s = 0; s < 32; ++s) {
      const float x = __uint_as_float(r[s]) + __uint_as_float((uint32_t)gxv[s] << 16);
      const float sg = __fdividef(1.0f, 1.0f + __expf(-ka * x));
      sAct[(gate * 32 + s) * 32 + unit] = fmaf(ka, sg, kb2);
    }
#pragma unroll
    for (int s = 0; s < 32; ++s) gxv[s] = gxn[s];
    __syncthreads();
    // cells (unit = lane, sequence 8 warp + x)
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const int s = warp * 8 + x;
      const int len = sLen[s];
      float h = 0.f;
      if (k < len) {
        const float gi = sAct[(0 * 32 + s) * 32 + lane], gf = sAct[(1 * 32 + s) * 32 + lane];
        const float gg = sAct[(2 * 32 + s) * 32 + lane], go = sAct[(3 * 32 + s) * 32 + lane];
        c_state[x] = gf * c_state[x] + gi * gg;
        h = go * tanhf_(c_state[x]);
        if (p.out_f32) {
          const int t = rev ? len - 1 - k : k;
          p.out_f32[(size_t)(sOff[s] + t) * p.ld_out_f32 + dir * 512 + (int)rank * 32 + lane] = h;
        }
      }
      const uint32_t j = piece_j0 + (uint32_t)(lane >> 3);
      *reinterpret_cast<__nv_bfloat16*>(sH + ((k + 1) & 1) * kTcH + piece_kb + s * 128 + ((j ^ (uint32_t)(s & 7)) << 4) + (lane & 7) * 2) =
          __float2bfloat16(h);
    }
    __syncthreads();
    // push this CTA's 2 KB piece (32 sequences x 4 chunks) to the peers; store the step's output rows
    {
      const int s = tid >> 2, c = tid & 3;
      const uint32_t j = piece_j0 + (uint32_t)c;
      const uint32_t loc = hn + piece_kb + s * 128 + ((j ^ (uint32_t)(s & 7)) << 4);
      const uint4 v = lds128(loc);
      if (!(p.dbg & 2)) {
#pragma unroll
        for (int d = 0; d < kCluster; ++d)
          if ((uint32_t)d != rank) st_cluster_128(mapa_u32(loc, (uint32_t)d), v);
      }
      const int len = sLen[s];
      if (p.out_bf16 && k < len) {
        const int t = rev ? len - 1 - k : k;
        stg128(p.out_bf16 + (size_t)(sOff[s] + t) * p.ld_out + dir * 512 + (int)rank * 32 + c * 8, v);
      }
    }
    if (!(p.dbg & 4)) asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes (local + remote) -> next step's tcgen05.mma
    if (!(p.dbg & 8)) cluster_sync_all(); else __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_d, 32);
  }
}
