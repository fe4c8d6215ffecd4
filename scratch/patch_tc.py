import re
p='/root/repo/coskad_b200/csrc/fused_eval_tc.cuh'
s=open(p).read()
orig_len=len(s)
# --- head weights vectorised
old='''            const float* wp = Pm.head_w + (c0 + u) * kP + p;
            float w[kDP];
#pragma unroll
            for (int d = 0; d < kDP; ++d) w[d] = __ldg(wp + d * kF);'''
new='''            const float4* wp = reinterpret_cast<const float4*>(Pm.head_w + (static_cast<size_t>(c0 + u) * kP + p) * kDP);
            float w[kDP];
#pragma unroll
            for (int d4 = 0; d4 < kDP / 4; ++d4) {
              const float4 t4 = __ldg(wp + d4);
              w[4 * d4 + 0] = t4.x; w[4 * d4 + 1] = t4.y; w[4 * d4 + 2] = t4.z; w[4 * d4 + 3] = t4.w;
            }'''
assert old in s; s=s.replace(old,new)
s=s.replace("const float* head_w;    // [16][kF]","const float* head_w;    // [kF][16]")
# --- constants / plans
a=s.index("constexpr int kTcThreads = 384;"); b=s.index("// blob = [B_hi image Kp*N]")
s=s[:a]+'''constexpr int kTcWarps = 12;                     // compute warps: 4 TMEM lane quarters x 3 groups
constexpr int kTcThreads = (kTcWarps + 1) * 32;  // + warp 12: dedicated tcgen05.mma issuer
constexpr int kTcTiles = 2 * kNW;                // M-tiles per CTA tile: (half j, window n) -> ti = j*kNW + n
// TMEM column plans (512 columns):
//   small phases (N <= 32: L1, L2, L3): D = 32 columns per M-tile at [0,192), five 64-column A buffers at [192,512)
//   big phases   (N = 64: the two halves of L4): D = 64 columns per M-tile at [0,384), two A buffers at [384,512)
// an A buffer = 32 hi + 32 lo columns (K <= 32 channels of one M-tile)
template <bool BIG> struct TcPlan {
  static constexpr uint32_t kDStride = BIG ? 64 : 32;
  static constexpr uint32_t kColA = BIG ? 384 : 192;
  __host__ __device__ static constexpr int buf(int ti) { return BIG ? (ti & 1) : (ti % 5); }
  __host__ __device__ static constexpr int use(int ti) { return BIG ? (ti >> 1) : (ti / 5); }
  __host__ __device__ static constexpr int uses(int b) { return BIG ? (b < 2 ? 3 : 0) : (b == 0 ? 2 : 1); }
};

'''+s[b:]
# --- pipe + phase + epilogue
a=s.index("struct TcPipe {"); b=s.index("__global__ void __launch_bounds__(kTcThreads, 1) fused_eval_tc_kernel")
new=open('/root/repo/scratch/tc_phase_new.txt').read()
s=s[:a]+new+s[b:]
# --- kernel body
old="                              + 16;                                     // mbarriers (5 x 8 B) + tmem base"
assert old in s; s=s.replace(old,"                              + 24;                                     // mbarriers (11 x 8 B) + tmem base")
old="  uint64_t* bars = reinterpret_cast<uint64_t*>(cen + 32);     // full[2], empty[2], done\n  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);"
assert old in s; s=s.replace(old,"  uint64_t* bars = reinterpret_cast<uint64_t*>(cen + 32);     // full[5], empty[5], done\n  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);")
old='''  if (tid == 32) {
    tc::mbar_init(&bars[0], 128); tc::mbar_init(&bars[1], 128);
    tc::mbar_init(&bars[2], 1);   tc::mbar_init(&bars[3], 1);
    tc::mbar_init(&bars[4], 1);
    tc::fence_mbar_init();
  }'''
assert old in s; s=s.replace(old,'''  if (tid == 32) {
    for (int b = 0; b < 5; ++b) { tc::mbar_init(&bars[b], 4); tc::mbar_init(&bars[5 + b], 1); }
    tc::mbar_init(&bars[10], 1);
    tc::fence_mbar_init();
  }''')
old='''  pipe.full = &bars[0]; pipe.empty = &bars[2]; pipe.done = &bars[4];
  pipe.tbase = *tmem_slot;
  pipe.n_full[0] = pipe.n_full[1] = 0; pipe.n_done = 0;
  // both A buffers start out free: one manual arrival completes phase 0 of the empty barriers
  if (tid == 0) { tc::mbar_arrive(&bars[2]); tc::mbar_arrive(&bars[3]); }
  pipe.n_empty[0] = pipe.n_empty[1] = 0;'''
assert old in s; s=s.replace(old,'''  pipe.full = &bars[0]; pipe.empty = &bars[5]; pipe.done = &bars[10];
  pipe.tbase = *tmem_slot;
  pipe.n_done = 0;
#pragma unroll
  for (int b = 0; b < 5; ++b) { pipe.n_full[b] = 0; pipe.n_empty[b] = 0; }
  // all A buffers start out free: one manual arrival completes phase 0 of every empty barrier
  if (tid == 0) { for (int b = 0; b < 5; ++b) tc::mbar_arrive(&bars[5 + b]); }''')
s=re.sub(r"(\n\s+)(temporal_stage<[^;]+;)", lambda m: m.group(1)+"if (warp < kTcWarps) "+m.group(2), s)
s=re.sub(r"(\n\s+)(spatial_stage<[^;]+;)", lambda m: m.group(1)+"if (warp < kTcWarps) "+m.group(2), s)
for o,n in (("tc_mix_phase<kC0, kC0, kC1>(","tc_mix_phase<kC0, kC0, kC1, false>("),("tc_mix_phase<kC1, 0, 2 * kC2>(","tc_mix_phase<kC1, 0, 2 * kC2, false>("),
            ("tc_mix_phase<kC2, kC2, kC3>(","tc_mix_phase<kC2, kC2, kC3, false>("),("tc_mix_phase<kC3, 0, kC4>(","tc_mix_phase<kC3, 0, kC4, true>(")):
    assert o in s; s=s.replace(o,n)
old="    {\n      const float* bias = WMb + 2 * 32 * 64;"
assert old in s; s=s.replace(old,"    if (warp < kTcWarps) {\n      const float* bias = WMb + 2 * 32 * 64;")
old="tc::tmem_ld16(pipe.tbase + (static_cast<uint32_t>(q * 32) << 16) + kTcColD + 64u * (j * kNW + n) + c0, v[n]);"
assert old in s; s=s.replace(old,"tc::tmem_ld16(pipe.tbase + (static_cast<uint32_t>(q * 32) << 16) + TcPlan<true>::kDStride * (j * kNW + n) + c0, v[n]);")
assert len(s) > orig_len
open(p,'w').write(s)
print('patched', len(s))
