p='/root/repo/coskad_b200/csrc/fused_eval_tc.cuh'
s=open(p).read()
a=s.index("      // 8 units (half j, 16-channel chunk) per lane quarter, split 3/3/2 over the three warp groups")
b=s.index("#pragma unroll\n      for (int n = 0; n < kNW; ++n)\n#pragma unroll\n        for (int d = 0; d < kDP; ++d) {\n          const float s = warp_sum(z[n][d]);")
new='''      // 8 units (half j, 16-channel chunk) per lane quarter, split 3/3/2 over the three warp groups.
      // The head weights W[(c*204+p)][16] stream L2 -> shared memory through a per-warp cp.async ring (kRing channels
      // deep) carved out of R0/R1, which are dead once layer 4 has been staged into TMEM: the L2 latency of this 278 KB
      // per window stream is hidden without holding prefetched weights in registers.
      constexpr int kRing = 6;
      float* ring = R0 + warp * (kRing * 512);                      // [slot][d4][lane][4]
      const int u0 = sub * 3, u1 = (sub == 2) ? 8 : u0 + 3;
      const int nco = (u1 - u0) * 16;                               // channels this warp visits, in unit order
      auto co_of = [&](int i, int& j, int& co) { const int unit = u0 + (i >> 4); j = unit >> 2; co = (unit & 3) * 16 + (i & 15); };
      auto prefetch = [&](int i) {
        if (i < nco) {
          int j, co; co_of(i, j, co);
          const int p = j * 128 + q * 32 + lane;
          if (p < kP) {
            const float* src = Pm.head_w + (static_cast<size_t>(co) * kP + p) * kDP;
            float* dst = ring + (i % kRing) * 512 + lane * 4;
#pragma unroll
            for (int d4 = 0; d4 < 4; ++d4) cp_async16(dst + d4 * 128, src + d4 * 4);
          }
        }
        cp_async_commit();
      };
#pragma unroll
      for (int i = 0; i < kRing - 1; ++i) prefetch(i);
      int i = 0;
      for (int unit = u0; unit < u1; ++unit) {
        const int j = unit >> 2, c0 = (unit & 3) * 16;
        const int p = j * 128 + q * 32 + lane;
        uint32_t v[kNW][16];
#pragma unroll
        for (int n = 0; n < kNW; ++n)
          tc::tmem_ld16(pipe.tbase + (static_cast<uint32_t>(q * 32) << 16) + TcPlan<true>::kDStride * (j * kNW + n) + c0, v[n]);
        tc::wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u, ++i) {
          prefetch(i + kRing - 1);
          asm volatile("cp.async.wait_group %0;\\n" ::"n"(kRing - 1) : "memory");     // channel i has landed (lane-private data)
          if (p < kP) {
            const float b = bias[c0 + u];
            float h[kNW];
#pragma unroll
            for (int n = 0; n < kNW; ++n) h[n] = prelu(__uint_as_float(v[n][u]) + b, slope4);
            const float4* wp = reinterpret_cast<const float4*>(ring + (i % kRing) * 512 + lane * 4);
            float w[kDP];
#pragma unroll
            for (int d4 = 0; d4 < kDP / 4; ++d4) {
              const float4 t4 = wp[d4 * 32];
              w[4 * d4 + 0] = t4.x; w[4 * d4 + 1] = t4.y; w[4 * d4 + 2] = t4.z; w[4 * d4 + 3] = t4.w;
            }
#pragma unroll
            for (int d = 0; d < kDP; ++d)
#pragma unroll
              for (int n = 0; n < kNW; ++n) z[n][d] = fmaf(h[n], w[d], z[n][d]);
          }
        }
      }
'''
s=s[:a]+new+s[b:]
open(p,'w').write(s)
print('ok')
