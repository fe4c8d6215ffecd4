p='/root/repo/coskad_b200/csrc/fused_eval_tc.cuh'
s=open(p).read()
def rep(o,n,cnt=1):
    global s
    assert s.count(o)>=1, o[:60]
    s=s.replace(o,n)
rep("constexpr uint32_t kTcColD = 0;                  // D: 64 columns per M-tile\nconstexpr uint32_t kTcColA = 64 * kTcTiles;      // A: 2 buffers x (32 hi + 32 lo) columns\nstatic_assert(kTcColA + 128 <= 512, \"TMEM budget\");",
'''// TMEM column plans (512 columns):
//   small phases (N <= 32: L1, L2, L3): D = 32 columns per M-tile at [0,192), five 64-column A buffers at [192,512)
//   big phases   (N = 64: the two halves of L4): D = 64 columns per M-tile at [0,384), two A buffers at [384,512)
// an A buffer = 32 hi + 32 lo columns (K <= 32 channels of one M-tile)
template <bool BIG> struct TcPlan {
  static constexpr uint32_t kDStride = BIG ? 64 : 32;
  static constexpr uint32_t kColA = BIG ? 384 : 192;
  __host__ __device__ static constexpr int buf(int ti) { return BIG ? (ti & 1) : (ti % 5); }
  __host__ __device__ static constexpr int use(int ti) { return BIG ? (ti >> 1) : (ti / 5); }
  __host__ __device__ static constexpr int uses(int b) { return BIG ? (b < 2 ? 3 : 0) : (b == 0 ? 2 : 1); }
};''')
rep("                              + 16;                                     // mbarriers (5 x 8 B) + tmem base","                              + 24;                                     // mbarriers (11 x 8 B) + tmem base")
rep('''  uint64_t* full;     // [2] A buffer b staged (128 arrivals)
  uint64_t* empty;    // [2] MMAs reading A buffer b complete (1 arrival: tcgen05.commit)''','''  uint64_t* full;     // [5] A buffer b staged (4 arrivals: lane 0 of each producer warp)
  uint64_t* empty;    // [5] MMAs reading A buffer b complete (1 arrival: tcgen05.commit)''')
rep("  uint32_t n_full[2], n_empty[2], n_done;","  uint32_t n_full[5], n_empty[5], n_done;")
rep("template <int K1, int K2, int N>\n__device__ __forceinline__ void tc_mix_phase(","template <int K1, int K2, int N, bool BIG>\n__device__ __forceinline__ void tc_mix_phase(")
rep("  static_assert(Kp <= 32 && N % 16 == 0 && N <= 64, \"bad mixing phase shape\");","  using Plan = TcPlan<BIG>;\n  static_assert(Kp <= 32 && N % 16 == 0 && N <= static_cast<int>(Plan::kDStride), \"bad mixing phase shape\");")
rep('''    const int b = sub;
    const uint32_t abuf = P.tbase + (static_cast<uint32_t>(q * 32) << 16) + kTcColA + 64u * b;
    for (int it = 0; it < kTcTiles / 2; ++it) {
      const int ti = b + 2 * it;''','''#pragma unroll
    for (int it = 0; it < kTcTiles / 2; ++it) {
      const int ti = sub + 2 * it;
      const int b = Plan::buf(ti);
      const uint32_t abuf = P.tbase + (static_cast<uint32_t>(q * 32) << 16) + Plan::kColA + 64u * b;''')
rep("      tc::mbar_wait(&P.empty[b], (P.n_empty[b] + it) & 1);","      tc::mbar_wait(&P.empty[b], (P.n_empty[b] + Plan::use(ti)) & 1);")
rep("      tc::fence_before_sync();\n      tc::mbar_arrive(&P.full[b]);","      tc::fence_before_sync();\n      __syncwarp();\n      if (lane == 0) tc::mbar_arrive(&P.full[b]);")
rep('''        const int b = ti & 1, it = ti >> 1;
        tc::mbar_wait(&P.full[b], (P.n_full[b] + it) & 1);''','''        const int b = Plan::buf(ti);
        tc::mbar_wait(&P.full[b], (P.n_full[b] + Plan::use(ti)) & 1);''')
rep('''        const uint32_t d = P.tbase + kTcColD + 64u * ti;
        const uint32_t a = P.tbase + kTcColA + 64u * b;''','''        const uint32_t d = P.tbase + Plan::kDStride * ti;
        const uint32_t a = P.tbase + Plan::kColA + 64u * b;''')
rep('''  P.n_full[0] += kTcTiles / 2; P.n_full[1] += kTcTiles / 2;
  P.n_empty[0] += kTcTiles / 2; P.n_empty[1] += kTcTiles / 2;''','''#pragma unroll
  for (int b = 0; b < 5; ++b) { P.n_full[b] += Plan::uses(b); P.n_empty[b] += Plan::uses(b); }''')
rep("    const uint32_t d = P.tbase + (static_cast<uint32_t>(q * 32) << 16) + kTcColD + 64u * ti;","    const uint32_t d = P.tbase + (static_cast<uint32_t>(q * 32) << 16) + TcPlan<false>::kDStride * ti;")
rep("  uint64_t* bars = reinterpret_cast<uint64_t*>(cen + 32);     // full[2], empty[2], done\n  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);","  uint64_t* bars = reinterpret_cast<uint64_t*>(cen + 32);     // full[5], empty[5], done\n  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);")
rep('''    tc::mbar_init(&bars[0], 128); tc::mbar_init(&bars[1], 128);
    tc::mbar_init(&bars[2], 1);   tc::mbar_init(&bars[3], 1);
    tc::mbar_init(&bars[4], 1);''','''    for (int b = 0; b < 5; ++b) { tc::mbar_init(&bars[b], 4); tc::mbar_init(&bars[5 + b], 1); }
    tc::mbar_init(&bars[10], 1);''')
rep('''  pipe.full = &bars[0]; pipe.empty = &bars[2]; pipe.done = &bars[4];
  pipe.tbase = *tmem_slot;
  pipe.n_full[0] = pipe.n_full[1] = 0; pipe.n_done = 0;
  // both A buffers start out free: one manual arrival completes phase 0 of the empty barriers
  if (tid == 0) { tc::mbar_arrive(&bars[2]); tc::mbar_arrive(&bars[3]); }
  pipe.n_empty[0] = pipe.n_empty[1] = 0;''','''  pipe.full = &bars[0]; pipe.empty = &bars[5]; pipe.done = &bars[10];
  pipe.tbase = *tmem_slot;
  pipe.n_done = 0;
#pragma unroll
  for (int b = 0; b < 5; ++b) { pipe.n_full[b] = 0; pipe.n_empty[b] = 0; }
  // all A buffers start out free: one manual arrival completes phase 0 of every empty barrier
  if (tid == 0) { for (int b = 0; b < 5; ++b) tc::mbar_arrive(&bars[5 + b]); }''')
for o,n in (("tc_mix_phase<kC0, kC0, kC1>(","tc_mix_phase<kC0, kC0, kC1, false>("),("tc_mix_phase<kC1, 0, 2 * kC2>(","tc_mix_phase<kC1, 0, 2 * kC2, false>("),
            ("tc_mix_phase<kC2, kC2, kC3>(","tc_mix_phase<kC2, kC2, kC3, false>("),("tc_mix_phase<kC3, 0, kC4>(","tc_mix_phase<kC3, 0, kC4, true>(")):
    rep(o,n)
rep("tc::tmem_ld16(pipe.tbase + (static_cast<uint32_t>(q * 32) << 16) + kTcColD + 64u * (j * kNW + n) + c0, v[n]);","tc::tmem_ld16(pipe.tbase + (static_cast<uint32_t>(q * 32) << 16) + TcPlan<true>::kDStride * (j * kNW + n) + c0, v[n]);")
open(p,'w').write(s)
print('ok')
