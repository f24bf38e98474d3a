// drsim_shard.cuh -- ONE persistent kernel for the step of clusters that do not fit a tile: a
// cluster of more than 1024 houses, and above all ONE very large cluster split by houses across
// the GPUs of a box (BASELINE config 5).
//
// It replaces the four launches k_house -> k_reduce -> k_env -> k_obs of the general path by one
// grid of resident CTAs (grid <= resident capacity, so every CTA is on an SM from the start); a CTA
// owns a contiguous run of 1024-house tiles:
//
//   phase 1  the CTA updates its tiles (hvac.py:43-64, building.py:141-222), keeps the post-update
//            state of its first tiles in shared memory and writes each tile's partial sums; then ONE
//            fence and one atomic per cluster it touched ARRIVE its tiles at the cluster's counter;
//   reduce   the CTA whose arrival completes cluster r combines the tile partials -- same chunk
//            geometry and combine order as k_house / k_reduce, so the sums are bit-identical to the
//            four-kernel path -- and, on a house-sharded cluster, stores this rank's row (and the halo
//            records of the shard's edge houses) straight into every peer's inbox over NVLink
//            (cluster.py:73-89 is the sum being distributed);
//   env      one thread waits (bounded) for the rows of all ranks, combines them in rank order, runs
//            the env epilogue (environment.py:87-106) and PUBLISHES the cluster's broadcast values
//            under a flag stamped with the exchange sequence number;
//   phase 2  every CTA waits for the flag of its tile's cluster and writes rewards and observation
//            rows (rewards_calculator.py:135-181, utils/norm.py:71-218).  PLAIN (fp32, 10-column rows
//            without neighbour messages -- the config-5 layout): every thread finishes the 4 houses it
//            updated, rows go through a whole-tile staging buffer and leave with one TMA bulk store per
//            warp.  Otherwise: one house per lane, 32-row groups, one TMA bulk store per group.
//
// Nothing crosses a kernel boundary inside the step: reduction, collective and the consumers of the
// reduced values are one launch.  Phase 1 never waits, so every arrival / push eventually happens and
// the waits of the env threads and of phase 2 cannot deadlock; all spins are bounded (~2 s) and
// report through ShardCtx::err / PeerCtx::err.
#pragma once

#include "drsim_kernels.cuh"

namespace drsim {

constexpr int kShardGroup = 32;    // rows per warp-level TMA store of the generic phase 2 (one row per lane)
constexpr int kShardMaxRuns = 64;  // (cluster, tile-count) runs a CTA arrives with one batch of atomics

struct ShardGeom {
  int chunks;       // 1024-house tiles per cluster: the chunk geometry of k_house (fixes the partial sums)
  int n_tiles;      // R * chunks
  int t_smem;       // tiles per CTA whose post-update state stays in shared memory between the phases
  int nbuf;         // generic phase 2: row-group staging buffers per warp (1 or 2)
  int tile_bytes;   // bytes of one saved tile
  int off_saved;    // [t_smem] saved tiles
  int off_rows;     // PLAIN: [1024][D] floats; generic: [warps][nbuf][kShardGroup][D] reals
  int smem_bytes;
};

struct ShardCtx {
  unsigned int *arrive;        // [R] tiles of cluster r that finished phase 1 this step (reset by the last arrival)
  unsigned long long *ready;   // [R] == StepIn::xseq once envb[r] holds this step's values
  void *envb;                  // [R][8] reals: EnvBroadcast of cluster r
  int *err;                    // set when a wait timed out
};

template <typename real>
struct ShardSaved {
  real *ta, *tm, *tg;
  int *sso;
  uint8_t *flags;
  DRSIM_D ShardSaved(unsigned char *base) {
    ta = reinterpret_cast<real *>(base);
    tm = ta + kTileSlots;
    tg = tm + kTileSlots;
    sso = reinterpret_cast<int *>(tg + kTileSlots);
    flags = reinterpret_cast<uint8_t *>(sso + kTileSlots);
  }
};

template <typename real>
DRSIM_D void shard_publish(const ShardCtx &sc, const StepIn &in, int r, const EnvBroadcast<real> &e) {
  real *dst = reinterpret_cast<real *>(sc.envb) + (size_t)r * 8;
  dst[0] = e.power_n; dst[1] = e.signal_n; dst[2] = e.solar_n; dst[3] = e.od_n;
  dst[4] = e.rew_sig; dst[5] = e.pen_common; dst[6] = e.pen_max;
  __threadfence();
#if defined(__CUDA_ARCH__)
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(sc.ready + r), "l"((unsigned long long)in.xseq) : "memory");
#endif
}

// thread 0: bounded wait for cluster r's broadcast values, copied into dst[8]
template <typename real>
DRSIM_D void shard_wait_env(const ShardCtx &sc, const StepIn &in, int r, real *dst) {
  unsigned long long v = 0;
  const long long t0 = clock64();
  for (;;) {
#if defined(__CUDA_ARCH__)
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(sc.ready + r) : "memory");
#endif
    if (v == (unsigned long long)in.xseq) break;
    if (clock64() - t0 > 4000000000ll) { *sc.err = 1; break; }   // ~2 s: give up, flag the error
  }
  const real *src = reinterpret_cast<const real *>(sc.envb) + (size_t)r * 8;
#pragma unroll
  for (int k = 0; k < 7; ++k) dst[k] = ld_cg(src + k);
}

DRSIM_D void ld4_cg(const float *p, float v[4]) {
#if defined(__CUDA_ARCH__)
  const float4 t = __ldcg(reinterpret_cast<const float4 *>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
#endif
}
DRSIM_D void ld4i_cg(const int32_t *p, int v[4]) {
#if defined(__CUDA_ARCH__)
  const int4 t = __ldcg(reinterpret_cast<const int4 *>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
#endif
}

// PLAIN = fp32 build, 10-column rows without neighbour messages (TarMAC layout / nb_comm = 0)
template <typename real, bool PLAIN>
__global__ void __launch_bounds__(kThreads, FusedOcc<real>::min_ctas)
k_shard(Planes<real> pl, SimParams p, StepIn in, ShardGeom g, ShardCtx sc, PeerCtx peer) {
  static_assert(!PLAIN || sizeof(real) == 4, "the plain variant is fp32 only");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_wp[kThreads / 32][kRed];
  __shared__ int s_run_r[kShardMaxRuns], s_run_n[kShardMaxRuns], s_run_last[kShardMaxRuns];
  __shared__ int s_n_runs, s_n_mine;
  __shared__ int s_mine[kShardMaxRuns];
  __shared__ real s_e[2][8];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ns = p.Ns, D = p.obs_dim;
  const int s0 = threadIdx.x * kHousesPerThread;
  const KC<real> kc(p);
  // this CTA's contiguous run of tiles (balanced: sizes differ by at most one)
  const int t_lo = (int)(((long long)blockIdx.x * g.n_tiles) / gridDim.x);
  const int t_hi = (int)(((long long)(blockIdx.x + 1) * g.n_tiles) / gridDim.x);
  if (threadIdx.x == 0) { s_n_runs = 0; s_n_mine = 0; }
  pdl_trigger();
  pdl_wait();

  // arrival of the runs recorded so far: one fence, one atomic per cluster (issued by different threads, so
  // their round trips overlap), then the reductions this CTA completed
  auto flush = [&]() {
    __threadfence();   // this thread's state (and partial) stores are visible device-wide before the arrivals below
    __syncthreads();
    const int nr = s_n_runs;
    if ((int)threadIdx.x < nr) {
      const int r = s_run_r[threadIdx.x], n = s_run_n[threadIdx.x];
      const unsigned old = atomicAdd(sc.arrive + r, (unsigned)n);
      const int last = old + (unsigned)n == (unsigned)g.chunks;
      if (last) sc.arrive[r] = 0;   // every tile of the cluster has arrived: ready for the next step
      __threadfence();
      s_run_last[threadIdx.x] = last;
    }
    __syncthreads();
    for (int k = 0; k < nr; ++k) {
      if (!s_run_last[k]) continue;   // CTA-uniform
      const int r = s_run_r[k];
      reduce_cluster<real>(pl, p, in, g.chunks, peer, r);
      if (threadIdx.x == 0) {
        if (peer.world > 1) s_mine[s_n_mine++] = r;   // the wait for the peers' rows is deferred: nothing left to contribute first
        else shard_publish<real>(sc, in, r, env_cluster<real>(pl, p, in, pl.acc, 1, peer, r));
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) s_n_runs = 0;
    if (peer.world > 1) {
      __syncthreads();
      const int nm = s_n_mine;
      if (lane == 0)
        for (int i = warp; i < nm; i += kThreads / 32) {
          const int r = s_mine[i];
          shard_publish<real>(sc, in, r, env_cluster<real>(pl, p, in, nullptr, 0, peer, r));
        }
      __syncthreads();
      if (threadIdx.x == 0) s_n_mine = 0;
    }
  };

  // ---- phase 1: house update of every tile of this CTA ------------------------------------------
  for (int tile = t_lo; tile < t_hi; ++tile) {
    const int it = tile - t_lo;
    const int r = tile / g.chunks, c = tile - r * g.chunks;
    const int n0 = c * kTileSlots + s0;
    real red[kRed] = {0, 0, 0, 0, 0};
    if (n0 < p.N) {
      House4<real> h;
      const size_t off = (size_t)r * Ns + n0;
      if constexpr (PLAIN) {
        Raw4f w;
        load4(pl.t_air + off, w.ta);
        load4(pl.t_mass + off, w.tm);
        load4i(pl.sso + off, w.sso);
        w.flags = load4b(pl.flags + off);
        load4_ro(pl.target + off, w.target);
        load4_ro(pl.cap + off, w.cap);
#pragma unroll
        for (int k = 0; k < 6; ++k) load4_ro(pl.coef[k] + off, w.c[k]);
        w.act = 0;
        if (p.policy == DRSIM_POLICY_EXTERNAL || p.policy == DRSIM_POLICY_GREEDY_MYOPIC)
          w.act = load4b((in.actions ? in.actions : pl.actions) + off);
        w.od = (float)pl.od_temp[r];
        w.solar = (float)pl.solar_next[r];
        house4_compute_f32<false>(pl, p, w, off, min(4, p.N - n0), h, red);
      } else {
        house4_step<real, true>(pl, p, kc, in, off, min(4, p.N - n0), (real)pl.od_temp[r], (real)pl.solar_next[r], h, red);
      }
      if (it < g.t_smem) {
        const ShardSaved<real> sv(smem_raw + g.off_saved + (size_t)it * g.tile_bytes);
        store4(sv.ta + s0, h.ta);
        store4(sv.tm + s0, h.tm);
        store4(sv.tg + s0, h.target);
        store4i(sv.sso + s0, h.sso);
        *reinterpret_cast<uint32_t *>(sv.flags + s0) = h.flags;
      }
    }
    double out[kRed];
    block_reduce<real, kThreads / 32>(red, out, s_wp);
    if (threadIdx.x == 0) {
      double *dst = pl.partials + ((size_t)r * g.chunks + c) * kRed;
      for (int k = 0; k < kRed; ++k) dst[k] = out[k];
      const int nr = s_n_runs;
      if (nr > 0 && s_run_r[nr - 1] == r) s_run_n[nr - 1]++;
      else { s_run_r[nr] = r; s_run_n[nr] = 1; s_n_runs = nr + 1; }
    }
    __syncthreads();   // s_wp is rewritten by the next tile; s_n_runs is read by everybody below
    if (s_n_runs == kShardMaxRuns && tile + 1 < t_hi) flush();
  }
  flush();

  // ---- phase 2: rewards + observation rows --------------------------------------------------------
  int ebuf = 0, r_have = -1;
  EnvBroadcast<real> e{};
  auto env_of = [&](int r) {   // CTA-uniform: (re)fetch the broadcast values when the cluster changes
    if (r == r_have) return;
    if (threadIdx.x == 0) shard_wait_env<real>(sc, in, r, s_e[ebuf]);
    __syncthreads();
    const real *v = s_e[ebuf];
    e.power_n = v[0]; e.signal_n = v[1]; e.solar_n = v[2]; e.od_n = v[3]; e.rew_sig = v[4]; e.pen_common = v[5]; e.pen_max = v[6];
    ebuf ^= 1;
    r_have = r;
  };

  if constexpr (PLAIN) {
    float *s_tile = reinterpret_cast<float *>(smem_raw + g.off_rows);
    bool store_pending = false;
    const int w0 = warp * 128;
    for (int tile = t_lo; tile < t_hi; ++tile) {
      const int it = tile - t_lo;
      const int r = tile / g.chunks, c = tile - r * g.chunks;
      env_of(r);
      const int n0 = c * kTileSlots + s0;
      const size_t rb = (size_t)r * Ns;
      const int slots = min(kTileSlots, Ns - c * kTileSlots);   // house slots of this tile (multiple of 4)
      // the warp's previous bulk store must have finished reading its rows before they are rewritten
      if (lane == 0 && store_pending) bulk_store_wait_read();
      __syncwarp();
      if (s0 < slots) {
        const size_t off = rb + n0;
        const int valid = min(4, p.N - n0);
        float ta[4], tm[4], tg[4];
        int sso[4];
        uint32_t flags;
        if (it < g.t_smem) {
          const ShardSaved<float> sv(smem_raw + g.off_saved + (size_t)it * g.tile_bytes);
          load4(sv.ta + s0, ta); load4(sv.tm + s0, tm); load4(sv.tg + s0, tg); load4i(sv.sso + s0, sso);
          flags = *reinterpret_cast<const uint32_t *>(sv.flags + s0);
        } else {
          ld4_cg(pl.t_air + off, ta); ld4_cg(pl.t_mass + off, tm); load4_ro(pl.target + off, tg); ld4i_cg(pl.sso + off, sso);
#if defined(__CUDA_ARCH__)
          flags = __ldcg(reinterpret_cast<const uint32_t *>(pl.flags + off));
#endif
        }
        float rw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = j < valid;
          rw[j] = ok ? house_reward<float>(p, kc, ta[j], tg[j], e) : 0.f;
          const uint32_t f = (flags >> (8 * j)) & 0xffu;
          const float sso_n = (float)fast_div((uint32_t)sso[j], p.fd_dur);          // norm.py:40-43, :79-82
          const float t20 = tg[j] - 20.f;
          float2 *r2 = reinterpret_cast<float2 *>(s_tile + (size_t)(s0 + j) * 10);
          r2[0] = make_float2(ok ? (float)(f & 1u) : 0.f, ok ? (float)((f >> 1) & 1u) : 0.f);
          r2[1] = make_float2(ok ? sso_n : 0.f, ok ? 1.f : 0.f);
          r2[2] = make_float2(ok ? e.power_n : 0.f, ok ? e.signal_n : 0.f);
          r2[3] = make_float2(ok ? p.hf.deadband : 0.f, ok ? (ta[j] + t20) * 0.2f : 0.f);
          r2[4] = make_float2(ok ? (tm[j] + t20) * 0.2f : 0.f, ok ? t20 * 0.2f : 0.f);
        }
        store4(pl.reward + off, rw);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && w0 < slots) {
        const int nrows = min(128, slots - w0);
        bulk_store_s2g(pl.obs + (rb + (size_t)c * kTileSlots + w0) * 10, s_tile + (size_t)w0 * 10, (uint32_t)(nrows * 40));
        store_pending = true;
      }
    }
    // shared memory must outlive the reads of the last row stores; their global writes complete with the grid
    if (lane == 0 && store_pending) bulk_store_wait_read();
  } else {
    real *rows_w = reinterpret_cast<real *>(smem_raw + g.off_rows) + (size_t)warp * g.nbuf * kShardGroup * D;
    const bool halo = needs_halo(p);
    int grp = 0;   // row groups this warp has shipped (selects the staging buffer)
    for (int tile = t_lo; tile < t_hi; ++tile) {
      const int it = tile - t_lo;
      const int r = tile / g.chunks, c = tile - r * g.chunks;
      env_of(r);
      const bool saved = it < g.t_smem;
      const ShardSaved<real> sv(smem_raw + g.off_saved + (size_t)(saved ? it : 0) * g.tile_bytes);
      const size_t rb = (size_t)r * Ns;
      for (int gi = 0; gi < 128 / kShardGroup; ++gi) {
        const int hs0 = warp * 128 + gi * kShardGroup;   // first house of the group inside the tile
        const int ng0 = c * kTileSlots + hs0;            // ... inside the cluster
        if (ng0 >= Ns) break;                            // warp-uniform: nothing of this group exists
        const int n = ng0 + lane, hs = hs0 + lane;
        real *row = rows_w + ((size_t)(grp % g.nbuf) * kShardGroup + lane) * D;
        if (D > 0) {
          // the bulk store that last used this staging buffer must have finished reading it
          if (lane == 0 && grp >= g.nbuf) {
#if defined(__CUDA_ARCH__)
            if (g.nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
          }
          __syncwarp();
        }
        if (n < p.N) {
          const size_t o = rb + n;
          real ta, tm, tgt;
          int sso;
          uint32_t f;
          if (saved) { ta = sv.ta[hs]; tm = sv.tm[hs]; tgt = sv.tg[hs]; sso = sv.sso[hs]; f = sv.flags[hs]; }
          else { ta = ld_cg(pl.t_air + o); tm = ld_cg(pl.t_mass + o); tgt = pl.target[o]; sso = ld_cg(pl.sso + o); f = ld_cg(pl.flags + o); }
          pl.reward[o] = house_reward<real>(p, kc, ta, tgt, e);
          if (D > 0) {
            real ratio[4] = {0, 0, 0, 0};
            if (p.st_thermal)
              for (int k = 0; k < 4; ++k) ratio[k] = pl.ratio[k][o];
            int i = obs_own<real>(row, p, f, (real)fast_div((uint32_t)sso, p.fd_dur), Rep<real>::minus20(ta, tgt),
                                  Rep<real>::minus20(tm, tgt), tgt - (real)20, e, ratio);
            if (p.obs_layout == DRSIM_OBS_HAND_ENGINEERED) {
              for (int k = 0; k < p.nb_comm; ++k) {
                int nb = neighbour_of(p, pl.comm_table, r, n + (halo ? (int)p.house_offset : 0), k);
                if (halo) {
                  // ring neighbour by GLOBAL index; outside this shard it comes from the exchanged halo
                  nb -= (int)p.house_offset;
                  if (nb < 0 || nb >= p.N) {
                    const int L = p.nb_comm / 2;
                    int d = nb < 0 ? nb + (int)p.n_global : nb;   // offset from the shard start, wrapped into [0, n_global)
                    d = d >= (int)p.n_global ? d - (int)p.n_global : d;
                    const bool left = d >= (int)p.n_global - L;
                    const double *rec = left ? in.halo_left + ((size_t)r * p.nb_comm + (d - ((int)p.n_global - L))) * kHaloFields
                                             : in.halo_right + ((size_t)r * p.nb_comm + (d - p.N)) * kHaloFields;
                    for (int m = 0; m < 4; ++m) row[i++] = (real)ld_cg(rec + m);
                    if (p.msg_thermal)
                      for (int m = 0; m < 4; ++m) row[i++] = (real)ld_cg(rec + 4 + m);
                    if (p.msg_hvac) { row[i++] = (real)p.cop; row[i++] = (real)p.latent; row[i++] = (real)p.dcap; }
                    continue;
                  }
                }
                const size_t q = rb + nb;
                const real pmax = qdiv(pl.cap[q], kc.cop, kc.inv_cop);
                row[i++] = div5(Rep<real>::dev(ld_cg(pl.t_air + q), pl.target[q]));       // norm.py:39
                row[i++] = (real)fast_div((uint32_t)ld_cg(pl.sso + q), p.fd_dur);         // norm.py:40-43
                row[i++] = qdiv((ld_cg(pl.flags + q) & 1u) ? pmax : (real)0, kc.nrs, kc.inv_nrs);
                row[i++] = qdiv(pmax, kc.nrs, kc.inv_nrs);
                if (p.msg_thermal)
                  for (int m = 0; m < 4; ++m) row[i++] = (real)pl.ratio[m][q];
                if (p.msg_hvac) {                                                         // constants, quirk Q11
                  row[i++] = (real)p.cop; row[i++] = (real)p.latent; row[i++] = (real)p.dcap;
                }
              }
            }
          }
        } else if (n < Ns) {
          for (int i = 0; i < D; ++i) row[i] = (real)0;  // padding slot
        }
        if (D > 0) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int nrows = min(kShardGroup, Ns - ng0);   // multiple of 4: the byte count is a multiple of 16
            bulk_store_s2g(pl.obs + (rb + ng0) * D, rows_w + (size_t)(grp % g.nbuf) * kShardGroup * D,
                           (uint32_t)((size_t)nrows * D * sizeof(real)));
          }
          ++grp;
        }
      }
    }
    if (lane == 0 && grp > 0) bulk_store_wait_read();
  }
}

}  // namespace drsim
