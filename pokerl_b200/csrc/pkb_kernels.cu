// pkb_kernels.cu -- CUDA kernels (sm_100a) and the C ABI of include/pokerl_b200.h.
//
// Kernels
//   step_kernel      K1+K2(+K5)  Game.step on a tile of envs staged in shared memory; end_hand and
//                                the ranked showdown run on block-compacted slot lists
//   reset_kernel     K3          Game.reset / setup_hand (dealing is a pure Philox function)
//   eval_*_kernel    K4          eval_hand on card lists / id masks / all C(52,k) hands
//   sample_kernel    K5          RandomAgent over the valid mask
//   (K6, the statistics reduction, is fused into step_kernel's epilogue)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/pokerl_b200.h"
#include "pkb_game.cuh"

namespace pkb {

// --------------------------------------------------------------------------------------------
// device-side state (one arena per handle)
// --------------------------------------------------------------------------------------------
struct DevState {
    double *cr, *bt, *pd, *mr, *payoffs;
    uint8_t* st;
    uint32_t *meta, *hand, *episode;
    unsigned long long* stats;  // [6] counters
    double* payoff_sum;         // [MAX_PLAYERS]
    uint64_t n, stride;
};

enum : int { STAT_STEPS = 0, STAT_HANDS, STAT_SHOWDOWNS, STAT_GAMES, STAT_INVALID, STAT_POISONED, STAT_COUNT };

struct StepArgs {
    DevState d;
    GameConst gc;
    const int32_t* actions;
    int32_t* actions_out;
    uint8_t* flags;
    uint8_t* valid;
    uint32_t step_idx;
    uint32_t use_weights;
    uint32_t weights[NUM_MOVES];
};

constexpr int TILE = 256;   // env-count granularity of the HBM planes (stride is a multiple of it)
constexpr int SLOTS = 96;   // per-warp shared-memory slots: 32 staging + 32 end-hand + 32 showdown
constexpr int STEP_WARPS = 4;  // warps per block of the step kernel (each warp is autonomous)

template <int P>
__host__ __device__ constexpr size_t warp_smem_bytes();

// resident blocks per SM the step kernel is compiled for: what 227 KB of shared memory allow, at most 7
template <int P>
__host__ __device__ constexpr int step_min_blocks()
{
    constexpr size_t per_block = warp_smem_bytes<P>() * STEP_WARPS + 1024;
    constexpr int by_smem = (int)((size_t)232448 / per_block);
#ifdef PKB_STEP_MAX_BLOCKS  // tuning knob for experiments
    constexpr int cap = PKB_STEP_MAX_BLOCKS;
#else
    constexpr int cap = 7;
#endif
    return by_smem > cap ? cap : (by_smem < 1 ? 1 : by_smem);
}

PKB_D View<0> global_view(const DevState& d)
{
    return View<0>{d.cr, d.bt, d.pd, d.mr, d.st, d.meta, d.hand, (size_t)d.stride};
}

// The warp's shared-memory slots plus the control words kept beside the game state, and the
// warp's statistics accumulators (written by lane 0 only)
template <int P>
struct WarpSlots {
    View<SLOTS> v;
    uint32_t* ctl;       // automaton control word (pkb_game.cuh, struct Ctl)
    uint32_t* env;       // local env index of the slot's occupant
    uint32_t* wstat;     // [STAT_COUNT] per-warp counters
    double* wpay;        // [P] per-warp payoff sums
};

template <int P>
__host__ __device__ constexpr size_t warp_smem_bytes()
{
    // per slot: f64 planes (3P+1), meta/hand/ctl/env u32, P state bytes; then P f64 + 8 u32 of statistics
    return (size_t)SLOTS * ((3 * P + 1) * 8 + 4 * 4 + P) + (size_t)P * 8 + 32;
}

template <int P>
PKB_D WarpSlots<P> carve_slots(unsigned char* base)
{
    WarpSlots<P> w;
    double* f64 = reinterpret_cast<double*>(base);
    w.v.cr = f64;
    w.v.bt = f64 + P * SLOTS;
    w.v.pd = f64 + 2 * P * SLOTS;
    w.v.mr = f64 + 3 * P * SLOTS;
    w.wpay = f64 + (3 * P + 1) * SLOTS;
    uint32_t* u32 = reinterpret_cast<uint32_t*>(f64 + (3 * P + 1) * SLOTS + P);
    w.v.meta = u32;
    w.v.hand = u32 + SLOTS;
    w.ctl = u32 + 2 * SLOTS;
    w.env = u32 + 3 * SLOTS;
    w.wstat = u32 + 4 * SLOTS;
    w.v.st = reinterpret_cast<uint8_t*>(u32 + 4 * SLOTS + 8);
    w.v.ld_rt = SLOTS;
    return w;
}

// warp-level event counter: lane 0 adds the number of lanes for which `event` holds
PKB_D void count_event(uint32_t* counter, bool event)
{
    const unsigned vote = __ballot_sync(0xffffffffu, event);
    if (vote && (threadIdx.x & 31) == 0) *counter += (uint32_t)__popc(vote);
}

template <int P>
PKB_D void move_slot(WarpSlots<P>& w, int from, int to)
{
#pragma unroll
    for (int p = 0; p < P; ++p) {
        w.v.cr[p * SLOTS + to] = w.v.cr[p * SLOTS + from];
        w.v.bt[p * SLOTS + to] = w.v.bt[p * SLOTS + from];
        w.v.pd[p * SLOTS + to] = w.v.pd[p * SLOTS + from];
        w.v.st[p * SLOTS + to] = w.v.st[p * SLOTS + from];
    }
    w.v.mr[to] = w.v.mr[from];
    w.v.meta[to] = w.v.meta[from];
    w.v.hand[to] = w.v.hand[from];
    w.ctl[to] = w.ctl[from];
    w.env[to] = w.env[from];
}

// ---- out-of-line pieces: ONE copy each in the kernel image.  The step kernel's warps run
//      different phases at the same time, so its code must stay small enough for the SM's
//      instruction cache; everything below takes and returns plain words.
__device__ __noinline__ uint32_t rank7_call(uint64_t mask) { return eval_mask52(mask); }

template <int P>
__device__ __noinline__ uint32_t setup_hand_call(unsigned char* warp_base, int slot, uint32_t meta_word, const GameConst& gc)
{
    WarpSlots<P> w = carve_slots<P>(warp_base);
    Meta m = unpack_meta(meta_word);
    setup_hand<P>(w.v, slot, m, gc);
    return pack_meta(m);
}

template <int P>
__device__ __noinline__ void showdown_call(unsigned char* warp_base, int slot, uint32_t ctl_word, uint64_t seed,
                                           uint32_t env_id, uint32_t episode)
{
    WarpSlots<P> w = carve_slots<P>(warp_base);
    showdown<P>(w.v, slot, unpack_ctl(ctl_word), seed, env_id, episode, [](uint64_t m) { return rank7_call(m); });
}

// Step epilogue for one finished env: flags, valid mask of the new active player, state back to HBM.
// Warp-uniform call; `active` selects the lanes that retire their slot.
template <int P>
PKB_D void retire_slot(WarpSlots<P>& w, int slot, bool active, const StepArgs& a)
{
    const DevState& d = a.d;
    Ctl c{PC_DONE, 0, 0, 0, 0, 0, 0, 0};
    if (active) c = unpack_ctl(w.ctl[slot]);
    const uint32_t f = (uint32_t)c.done | (c.invalid ? F_INVALID : 0u) | (c.poisoned ? F_POISONED : 0u);
    if (a.gc.track_stats) {
        count_event(&w.wstat[STAT_STEPS], active && !c.invalid);
        count_event(&w.wstat[STAT_GAMES], active && (f & F_GAME_OVER));
        count_event(&w.wstat[STAT_INVALID], active && c.invalid);
        count_event(&w.wstat[STAT_POISONED], active && c.poisoned);
    }
    if (!active) return;
    const size_t env = w.env[slot];
    if (a.flags) a.flags[env] = (uint8_t)f;
    if (a.valid) a.valid[env] = (uint8_t)valid_mask<P>(w.v, slot, unpack_meta(w.v.meta[slot]).act);
    if (c.invalid) return;  // nothing changed (game.py:649-651 raises before any mutation)
#pragma unroll
    for (int p = 0; p < P; ++p) {
        d.cr[p * d.stride + env] = w.v.cr[p * SLOTS + slot];
        d.bt[p * d.stride + env] = w.v.bt[p * SLOTS + slot];
        d.pd[p * d.stride + env] = w.v.pd[p * SLOTS + slot];
        d.st[p * d.stride + env] = w.v.st[p * SLOTS + slot];
    }
    d.mr[env] = w.v.mr[slot];
    d.meta[env] = w.v.meta[slot];
    d.hand[env] = w.v.hand[slot];
}

// What follows Game.end_hand once the payoffs are in the pending plane: finish the hand, deal the next
// one, resume the seat scan -- and any further hand that ends inside the same step (rare: all-in blinds
// make whole hands play themselves) -- then auto-reset if asked.  Warp-uniform, out of line.
// `ctl_word` must have pc == PC_END_HAND semantics already resolved by end_hand_begin (+ showdown);
// `asserted`: end_hand_begin found no potential winner (game.py:473 would raise AssertionError).
template <int P>
__device__ __noinline__ uint32_t finish_hand_call(unsigned char* warp_base, int slot, uint32_t ctl_word, bool active,
                                                  bool asserted, const StepArgs& a)
{
    WarpSlots<P> w = carve_slots<P>(warp_base);
    const DevState& d = a.d;
    Ctl c = unpack_ctl(ctl_word);
    double paysum[P];
#pragma unroll
    for (int p = 0; p < P; ++p) paysum[p] = 0.0;
    uint32_t more_hands = 0, more_shows = 0;
    if (active) {
        Meta m = unpack_meta(w.v.meta[slot]);
        const size_t env = w.env[slot];
        const auto deal = [&](const Meta& mm) { return unpack_meta(setup_hand_call<P>(warp_base, slot, pack_meta(mm), a.gc)); };
        while (!asserted) {
            double pay[P];
            end_hand_finish<P>(w.v, slot, m, c, d.payoffs, d.stride, env, pay, deal);
#pragma unroll
            for (int p = 0; p < P; ++p) paysum[p] += pay[p];
            if (c.pc != PC_END_HAND) break;
            ++more_hands;  // another hand ended inside this step
            if (end_hand_begin<P>(w.v, slot, c)) {
                ++more_shows;
                showdown_call<P>(warp_base, slot, pack_ctl(c), a.gc.seed, a.gc.env_id_base + (uint32_t)env, d.episode[env]);
            } else if (c.poisoned)
                asserted = true;
        }
        if (asserted) {  // payoffs were zeroed at game.py:468 before the assert
#pragma unroll
            for (int p = 0; p < P; ++p) d.payoffs[p * d.stride + env] = 0.0;
            c.done = c.cont ? (int)((game_over<P>(w.v, slot) ? F_GAME_OVER : 0u) | F_HAND_OVER | F_ROUND_OVER)
                            : (int)F_HAND_OVER;
            c.pc = PC_DONE;
        } else if ((c.done & (int)F_GAME_OVER) && !c.poisoned && a.gc.auto_reset) {
            // Game.reset() right after the step that ended the game (game.py:397-412)
            d.episode[env] += 1;
            m.dealer = a.gc.dealer;
            m.act = 0;
            w.v.hand[slot] = 0;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                w.v.cr[p * SLOTS + slot] = a.gc.start_credits[p];
                w.v.st[p * SLOTS + slot] = ST_ACTIVE;
            }
            m = deal(m);
        }
        w.v.meta[slot] = pack_meta(m);
    }
    if (a.gc.track_stats) {
        const uint32_t h = __reduce_add_sync(0xffffffffu, more_hands), sh = __reduce_add_sync(0xffffffffu, more_shows);
        if ((threadIdx.x & 31) == 0) {
            w.wstat[STAT_HANDS] += h;
            w.wstat[STAT_SHOWDOWNS] += sh;
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            double x = paysum[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) == 0) w.wpay[p] += x;
        }
    }
    return pack_ctl(c);
}

// Ranked showdowns of the warp's showdown buffer (slots 64..64+count), one env per lane
template <int P>
__device__ __noinline__ void flush_showdowns(unsigned char* warp_base, int s_count, const StepArgs& a)
{
    WarpSlots<P> w = carve_slots<P>(warp_base);
    const int lane = threadIdx.x & 31, slot = 64 + lane;
    const bool active = lane < s_count;
    __syncwarp();
    uint32_t ctl_word = 0;
    if (active) {
        ctl_word = w.ctl[slot];
        const size_t env = w.env[slot];
        showdown_call<P>(warp_base, slot, ctl_word, a.gc.seed, a.gc.env_id_base + (uint32_t)env, a.d.episode[env]);
    }
    if (a.gc.track_stats) count_event(&w.wstat[STAT_SHOWDOWNS], active);
    ctl_word = finish_hand_call<P>(warp_base, slot, ctl_word, active, false, a);
    if (active) w.ctl[slot] = ctl_word;
    retire_slot<P>(w, slot, active, a);
    __syncwarp();
}

// end_hand for the warp's end-hand buffer (slots 32..32+count): last-stand hands finish here,
// ranked hands move on to the showdown buffer.  Returns the new fill of the showdown buffer.
template <int P>
__device__ __noinline__ int flush_end_hands(unsigned char* warp_base, int e_count, int s_count, const StepArgs& a)
{
    WarpSlots<P> w = carve_slots<P>(warp_base);
    const int lane = threadIdx.x & 31, slot = 32 + lane;
    const bool active = lane < e_count;
    __syncwarp();
    bool ranked = false;
    uint32_t ctl_word = 0;
    bool asserted = false;
    if (active) {
        Ctl c = unpack_ctl(w.ctl[slot]);
        ranked = end_hand_begin<P>(w.v, slot, c);
        asserted = !ranked && c.poisoned;
        if (ranked) c.pc = PC_SHOWDOWN;
        ctl_word = pack_ctl(c);
        w.ctl[slot] = ctl_word;
    }
    if (a.gc.track_stats) count_event(&w.wstat[STAT_HANDS], active);
    ctl_word = finish_hand_call<P>(warp_base, slot, ctl_word, active && !ranked, asserted, a);
    if (active && !ranked) w.ctl[slot] = ctl_word;
    retire_slot<P>(w, slot, active && !ranked, a);
    const unsigned vote = __ballot_sync(0xffffffffu, ranked);
    const int k = __popc(vote);
    if (k) {
        if (s_count + k > 32) {
            flush_showdowns<P>(warp_base, s_count, a);
            s_count = 0;
        }
        __syncwarp();
        if (ranked) move_slot<P>(w, slot, 64 + s_count + __popc(vote & ((1u << lane) - 1u)));
        s_count += k;
        if (s_count == 32) {
            flush_showdowns<P>(warp_base, s_count, a);
            s_count = 0;
        }
    }
    __syncwarp();
    return s_count;
}

// --------------------------------------------------------------------------------------------
// K1/K2: Game.step (game.py:621-700).  Persistent grid; every warp owns 96 shared-memory slots
// and walks 32-env sub-tiles on its own: phase A (validate, bet, scan seats) runs one env per lane
// on the staging slots; envs whose hand ends are parked in the warp's end-hand buffer and those
// that need ranking in its showdown buffer, and each buffer is processed when it fills, so the
// expensive branches run at (nearly) full warp width without any block-level barrier.
// --------------------------------------------------------------------------------------------
template <int P, bool SAMPLE>
__global__ void __launch_bounds__(STEP_WARPS * 32, step_min_blocks<P>()) step_kernel(const __grid_constant__ StepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char* warp_base = smem_raw + (size_t)wib * warp_smem_bytes<P>();
    WarpSlots<P> w = carve_slots<P>(warp_base);
    const DevState& d = a.d;
    if (lane < 8) w.wstat[lane] = 0;
    if (lane < P) w.wpay[lane] = 0.0;
    int e_count = 0, s_count = 0;  // warp-uniform fill of the end-hand / showdown buffers
    __syncwarp();

    const size_t n_sub = d.stride / 32, total_warps = (size_t)gridDim.x * STEP_WARPS;
    for (size_t sub = (size_t)blockIdx.x * STEP_WARPS + wib; sub < n_sub; sub += total_warps) {
        const size_t env = sub * 32 + lane;
        const bool live = env < d.n;
        // ---- stage 32 envs: HBM planes -> staging slots (slot = lane; 256 B per plane row)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            w.v.cr[p * SLOTS + lane] = d.cr[p * d.stride + env];
            w.v.bt[p * SLOTS + lane] = d.bt[p * d.stride + env];
            w.v.pd[p * SLOTS + lane] = d.pd[p * d.stride + env];
            w.v.st[p * SLOTS + lane] = d.st[p * d.stride + env];
        }
        w.v.mr[lane] = d.mr[env];
        w.v.meta[lane] = d.meta[env];
        w.v.hand[lane] = d.hand[env];
        w.env[lane] = (uint32_t)env;

        // ---- phase A: validate the action, bet, scan seats up to the first end_hand
        Ctl c{PC_DONE, 0, 0, 0, 0, 0, 0, 0};
        if (live) {
            Meta m = unpack_meta(w.v.meta[lane]);
            const uint32_t vm = valid_mask<P>(w.v, lane, m.act);
            int action;
            if (SAMPLE) {
                action = sample_action(a.gc.seed, a.gc.env_id_base + (uint32_t)env, a.step_idx, vm,
                                       a.use_weights ? a.weights : nullptr);
                if (a.actions_out) a.actions_out[env] = action;
            } else
                action = a.actions[env];
            step_begin<P>(w.v, lane, m, c, action, vm, a.gc);
            w.v.meta[lane] = pack_meta(m);
        }
        w.ctl[lane] = pack_ctl(c);
        retire_slot<P>(w, lane, live && c.pc == PC_DONE, a);

        // ---- park the envs whose hand ended; process the buffer when it is full
        const bool parked = live && c.pc == PC_END_HAND;
        const unsigned vote = __ballot_sync(0xffffffffu, parked);
        const int k = __popc(vote);
        if (k) {
            if (e_count + k > 32) {
                s_count = flush_end_hands<P>(warp_base, e_count, s_count, a);
                e_count = 0;
            }
            __syncwarp();
            if (parked) move_slot<P>(w, lane, 32 + e_count + __popc(vote & ((1u << lane) - 1u)));
            e_count += k;
            if (e_count == 32) {
                s_count = flush_end_hands<P>(warp_base, e_count, s_count, a);
                e_count = 0;
            }
        }
        __syncwarp();
    }
    // ---- drain what is left in the buffers
    if (e_count) s_count = flush_end_hands<P>(warp_base, e_count, s_count, a);
    if (s_count) flush_showdowns<P>(warp_base, s_count, a);

    // ---- K6: statistics, one set of global atomics per warp for the whole launch
    __syncwarp();
    if (a.gc.track_stats) {
        if (lane < STAT_COUNT && w.wstat[lane]) atomicAdd(&d.stats[lane], (unsigned long long)w.wstat[lane]);
        if (lane < P && w.wpay[lane] != 0.0) atomicAdd(&d.payoff_sum[lane], w.wpay[lane]);
    }
}

// --------------------------------------------------------------------------------------------
// K3: Game.reset (game.py:397-412), optionally masked
// --------------------------------------------------------------------------------------------
template <int P>
__global__ void reset_kernel(DevState d, GameConst gc, const uint8_t* mask)
{
    const size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n || (mask && !mask[env])) return;
    View<0> v = global_view(d);
    d.episode[env] += 1;
    Meta m = unpack_meta(d.meta[env]);
    reset_env<P>(v, env, m, gc);
    d.meta[env] = pack_meta(m);
}

template <int P>
__global__ void valid_kernel(DevState d, int player, uint8_t* out)
{
    const size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n) return;
    const View<0> v = global_view(d);
    const int who = player < 0 ? unpack_meta(d.meta[env]).act : player;
    out[env] = (uint8_t)valid_mask<P>(v, env, who);
}

template <int P>
__global__ void game_over_kernel(DevState d, uint8_t* out)
{
    const size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n) return;
    out[env] = game_over<P>(global_view(d), env) ? 1 : 0;
}

// K5: RandomAgent (agents/random.py:16-17) / weighted mix over the valid mask
template <int P>
__global__ void sample_kernel(DevState d, GameConst gc, uint32_t step_idx, uint32_t use_weights, const uint32_t* weights,
                              int32_t* actions)
{
    __shared__ uint32_t w[NUM_MOVES];
    if (threadIdx.x < NUM_MOVES) w[threadIdx.x] = use_weights ? weights[threadIdx.x] : 1u;
    __syncthreads();
    const size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n) return;
    const View<0> v = global_view(d);
    const uint32_t vm = valid_mask<P>(v, env, unpack_meta(d.meta[env]).act);
    actions[env] = sample_action(gc.seed, gc.env_id_base + (uint32_t)env, step_idx, vm, w);
}

// The current deal of every env as card values: rows 0..4 board, 5+2p / 6+2p hole cards of seat p
template <int P>
__global__ void cards_kernel(DevState d, GameConst gc, uint8_t* out)
{
    const size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n) return;
    Dealer dl(gc.seed, gc.env_id_base + (uint32_t)env, d.episode[env], d.hand[env]);
#pragma unroll
    for (int i = 0; i < 5 + 2 * P; ++i) out[(size_t)i * d.stride + env] = (uint8_t)card_value_from_id(dl.next());
}

// eval_hand(StateView.player_hand) (game.py:139-143; q_learning.py:29-33): hole cards of `player`
// plus the community cards visible on the current street (game.py:278)
template <int P>
__global__ void features_kernel(DevState d, GameConst gc, int player, uint32_t* out)
{
    const size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n) return;
    const Meta m = unpack_meta(d.meta[env]);
    const int who = player < 0 ? m.act : player;
    const int shown = m.turn == 0 ? 0 : m.turn + 2;
    Dealer dl(gc.seed, gc.env_id_base + (uint32_t)env, d.episode[env], d.hand[env]);
    uint64_t mask = 0;
#pragma unroll
    for (int i = 0; i < 5 + 2 * P; ++i) {
        const uint32_t id = dl.next();
        const bool take = i < 5 ? i < shown : (i - 5) >> 1 == who;
        if (take) mask |= 1ull << id;
    }
    out[env] = eval_mask52(mask);
}

// --------------------------------------------------------------------------------------------
// K4: judger
// --------------------------------------------------------------------------------------------
__global__ void eval_cards_kernel(const uint8_t* __restrict__ cards, int k, uint64_t n, uint32_t* __restrict__ out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t c[7];
        for (int j = 0; j < k; ++j) c[j] = cards[i * k + j];
        out[i] = eval_cards(c, k);
    }
}

__global__ void eval_masks_kernel(const uint64_t* __restrict__ masks, uint64_t n, uint32_t* __restrict__ out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = eval_mask52(masks[i]);
}

__constant__ uint64_t c_binom[53][8];  // C(n, k) for n <= 52, k <= 7

// index -> k-subset of {0..51} in colex order (combinatorial number system)
PKB_D uint64_t unrank_colex(uint64_t idx, int k)
{
    uint64_t mask = 0;
    int hi = 52;
    for (int i = k; i >= 1; --i) {
        // largest c < hi with C(c, i) <= idx
        int lo = i - 1, up = hi - 1;
        while (lo < up) {
            const int mid = (lo + up + 1) >> 1;
            if (c_binom[mid][i] <= idx) lo = mid;
            else up = mid - 1;
        }
        mask |= 1ull << lo;
        idx -= c_binom[lo][i];
        hi = lo;
    }
    return mask;
}

constexpr int EXH_CHUNK = 32;  // hands per thread; 6-bit category counters stay below 64

// all C(52,k) hands, enumerated on chip: 0 bytes in, 4 bytes out (optional)
__global__ void __launch_bounds__(256)
eval_exhaustive_kernel(int k, uint64_t first, uint64_t count, uint32_t* __restrict__ out, unsigned long long* hist,
                       unsigned long long* sums)
{
    __shared__ unsigned long long s_hist[11], s_sum[2];
    if (threadIdx.x < 11) s_hist[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_sum[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = t * EXH_CHUNK;
    unsigned long long cats = 0, sum_a = 0, sum_b = 0;
    if (lo < count) {
        const int todo = (int)(count - lo < EXH_CHUNK ? count - lo : EXH_CHUNK);
        uint64_t m = unrank_colex(first + lo, k);
        for (int j = 0; j < todo; ++j) {
            const uint32_t p = eval_mask52(m);
            if (out) out[lo + j] = p;
            cats += 1ull << (6 * (p >> 20));
            sum_a += p;
            sum_b += (unsigned long long)p * (m % 1000003ull);
            // next mask with the same popcount (Gosper), the division by a power of two as a shift
            const uint64_t c = m & (0 - m), r = m + c;
            m = (((r ^ m) >> 2) >> (__ffsll((long long)c) - 1)) | r;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_a += __shfl_xor_sync(0xffffffffu, sum_a, o);
        sum_b += __shfl_xor_sync(0xffffffffu, sum_b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_sum[0], sum_a);
        atomicAdd(&s_sum[1], sum_b);
    }
    for (int cat = 1; cat <= 10; ++cat) {
        const uint32_t mine = (uint32_t)(cats >> (6 * cat)) & 63u;
        const uint32_t w = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_hist[cat], (unsigned long long)w);
    }
    __syncthreads();
    if (threadIdx.x < 11 && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
    if (threadIdx.x < 2) atomicAdd(&sums[threadIdx.x], s_sum[threadIdx.x]);
}

// n random 7-card hands: counter {index_lo, index_hi, block, 0x40000000}
__global__ void random_hands_kernel(uint64_t seed, uint64_t n, uint64_t* __restrict__ masks)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t free_mask = (1ull << 52) - 1;
        U4 b = philox4x32_10(U4{(uint32_t)i, (uint32_t)(i >> 32), 0u, 0x40000000u}, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t w0[4] = {b.x, b.y, b.z, b.w};
        b = philox4x32_10(U4{(uint32_t)i, (uint32_t)(i >> 32), 1u, 0x40000000u}, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t w1[3] = {b.x, b.y, b.z};
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const uint32_t word = j < 4 ? w0[j] : w1[j - 4];
            free_mask &= ~(1ull << select_bit(free_mask, mulhi32(word, (uint32_t)(52 - j))));
        }
        masks[i] = ~free_mask & ((1ull << 52) - 1);
    }
}

// compare_rankings (judger.py:111-158) on n tables of p rankings
__global__ void compare_kernel(const uint32_t* __restrict__ packed, int p, uint64_t n, uint8_t* __restrict__ onehot)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        WinnerScan ws;
        for (int j = 0; j < p; ++j) ws.feed(j, packed[i * p + j]);
        for (int j = 0; j < p; ++j) onehot[i * p + j] = (uint8_t)(ws.winners >> j & 1u);
    }
}

}  // namespace pkb

// ================================================================================================
// C ABI
// ================================================================================================
using namespace pkb;

struct pkb_env {
    pkb_config cfg;
    DevState d;
    GameConst gc;
    void* arena;
    size_t arena_bytes;
    uint32_t* d_weights;
};

static thread_local std::string g_error;

static int fail(const char* what, cudaError_t e = cudaSuccess)
{
    g_error = what;
    if (e != cudaSuccess) {
        g_error += ": ";
        g_error += cudaGetErrorString(e);
    }
    return -1;
}

#define PKB_CUDA(call)                                  \
    do {                                                \
        cudaError_t e_ = (call);                        \
        if (e_ != cudaSuccess) return fail(#call, e_);  \
    } while (0)

#define PKB_DISPATCH_P(P_, ...)                                         \
    switch (P_) {                                                       \
        case 2: { constexpr int P = 2; __VA_ARGS__; } break;            \
        case 3: { constexpr int P = 3; __VA_ARGS__; } break;            \
        case 4: { constexpr int P = 4; __VA_ARGS__; } break;            \
        case 5: { constexpr int P = 5; __VA_ARGS__; } break;            \
        case 6: { constexpr int P = 6; __VA_ARGS__; } break;            \
        case 7: { constexpr int P = 7; __VA_ARGS__; } break;            \
        case 8: { constexpr int P = 8; __VA_ARGS__; } break;            \
        case 9: { constexpr int P = 9; __VA_ARGS__; } break;            \
        case 10: { constexpr int P = 10; __VA_ARGS__; } break;          \
        default: return fail("num_players must be in 2..10");           \
    }

static inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

template <int P, bool SAMPLE>
static cudaError_t launch_step(const StepArgs& a, cudaStream_t s)
{
    constexpr size_t smem = warp_smem_bytes<P>() * STEP_WARPS;
    static int grid_blocks[64] = {0};  // per device: persistent grid = SMs x resident blocks per SM
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) dev = 0;
    if (!grid_blocks[dev]) {
        e = cudaFuncSetAttribute(step_kernel<P, SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel<P, SAMPLE>, STEP_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        grid_blocks[dev] = (per_sm > 0 ? per_sm : 1) * sms;
    }
    const size_t n_sub = a.d.stride / 32, want = (n_sub + STEP_WARPS - 1) / STEP_WARPS;
    const unsigned grid = (unsigned)(want < (size_t)grid_blocks[dev] ? want : (size_t)grid_blocks[dev]);
    step_kernel<P, SAMPLE><<<grid, STEP_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

extern "C" {

const char* pkb_last_error(void) { return g_error.c_str(); }
int pkb_abi_version(void) { return PKB_ABI_VERSION; }

int pkb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fail("cudaGetDeviceCount", e);
        return -1;
    }
    return n;
}

int pkb_create(const pkb_config* cfg, pkb_env** out)
{
    if (!cfg || !out) return fail("pkb_create: null argument");
    if (cfg->abi_version != PKB_ABI_VERSION) return fail("pkb_create: ABI version mismatch");
    if (cfg->num_players < 2 || cfg->num_players > PKB_MAX_PLAYERS) return fail("pkb_create: num_players must be in 2..10");
    if (cfg->num_envs == 0) return fail("pkb_create: num_envs must be > 0");
    if (cfg->env_id_base + cfg->num_envs > (1ull << 32)) return fail("pkb_create: global env ids must be < 2^32");
    if (cfg->dealer < 0 || cfg->dealer >= cfg->num_players) return fail("pkb_create: dealer out of range");
    PKB_CUDA(cudaSetDevice(cfg->device));

    pkb_env* h = new pkb_env();
    h->cfg = *cfg;
    const int P = cfg->num_players;
    const uint64_t n = cfg->num_envs, stride = (n + TILE - 1) / TILE * TILE;
    // arena: f64 planes | u32 planes | u8 planes | stats
    const size_t n_f64 = (size_t)(4 * P + 1) * stride, n_u32 = (size_t)3 * stride, n_u8 = (size_t)P * stride;
    const size_t off_u32 = n_f64 * 8, off_u8 = off_u32 + n_u32 * 4, off_stats = (off_u8 + n_u8 + 255) / 256 * 256;
    h->arena_bytes = off_stats + STAT_COUNT * 8 + MAX_PLAYERS * 8 + NUM_MOVES * 4;
    cudaError_t e = cudaMalloc(&h->arena, h->arena_bytes);
    if (e != cudaSuccess) {
        delete h;
        return fail("pkb_create: cudaMalloc", e);
    }
    e = cudaMemset(h->arena, 0, h->arena_bytes);
    if (e != cudaSuccess) {
        cudaFree(h->arena);
        delete h;
        return fail("pkb_create: cudaMemset", e);
    }
    char* base = static_cast<char*>(h->arena);
    double* f = reinterpret_cast<double*>(base);
    DevState& d = h->d;
    d.cr = f;
    d.bt = f + (size_t)P * stride;
    d.pd = f + (size_t)2 * P * stride;
    d.payoffs = f + (size_t)3 * P * stride;
    d.mr = f + (size_t)4 * P * stride;
    uint32_t* u = reinterpret_cast<uint32_t*>(base + off_u32);
    d.meta = u;
    d.hand = u + stride;
    d.episode = u + 2 * stride;
    d.st = reinterpret_cast<uint8_t*>(base + off_u8);
    d.stats = reinterpret_cast<unsigned long long*>(base + off_stats);
    d.payoff_sum = reinterpret_cast<double*>(base + off_stats + STAT_COUNT * 8);
    h->d_weights = reinterpret_cast<uint32_t*>(base + off_stats + STAT_COUNT * 8 + MAX_PLAYERS * 8);
    d.n = n;
    d.stride = stride;

    GameConst& gc = h->gc;
    memset(&gc, 0, sizeof gc);
    for (int p = 0; p < P; ++p) gc.start_credits[p] = cfg->start_credits[p];
    gc.big_blind = cfg->big_blind;
    gc.small_blind = cfg->small_blind;
    gc.seed = cfg->seed;
    gc.env_id_base = (uint32_t)cfg->env_id_base;
    gc.dealer = cfg->dealer;
    gc.auto_reset = cfg->auto_reset;
    gc.track_stats = 1;
    *out = h;
    return 0;
}

int pkb_destroy(pkb_env* h)
{
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->arena);
    delete h;
    return 0;
}

int pkb_set_reset_dealer(pkb_env* h, int32_t dealer)
{
    if (!h) return fail("pkb_set_reset_dealer: null handle");
    if (dealer < 0 || dealer >= h->cfg.num_players) return fail("pkb_set_reset_dealer: dealer out of range");
    h->cfg.dealer = dealer;
    h->gc.dealer = dealer;
    return 0;
}

int pkb_reset(pkb_env* h, const uint8_t* env_mask, void* stream)
{
    if (!h) return fail("pkb_reset: null handle");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_DISPATCH_P(h->cfg.num_players, (reset_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, h->gc, env_mask)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

static int step_common(pkb_env* h, bool sample, uint32_t step_idx, const uint32_t* weights, const int32_t* actions,
                       int32_t* actions_out, uint8_t* flags, uint8_t* valid, void* stream)
{
    if (!h) return fail("pkb_step: null handle");
    if (!sample && !actions) return fail("pkb_step: null actions");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    StepArgs a;
    a.d = h->d;
    a.gc = h->gc;
    a.actions = actions;
    a.actions_out = actions_out;
    a.flags = flags;
    a.valid = valid;
    a.step_idx = step_idx;
    a.use_weights = weights ? 1u : 0u;
    for (int i = 0; i < NUM_MOVES; ++i) a.weights[i] = weights ? weights[i] : 1u;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    PKB_DISPATCH_P(h->cfg.num_players,
                   e = sample ? launch_step<P, true>(a, s) : launch_step<P, false>(a, s));
    if (e != cudaSuccess) return fail("pkb_step: launch", e);
    return 0;
}

int pkb_step(pkb_env* h, const int32_t* actions, uint8_t* flags, uint8_t* valid, void* stream)
{
    return step_common(h, false, 0, nullptr, actions, nullptr, flags, valid, stream);
}

int pkb_step_random(pkb_env* h, uint32_t step_idx, const uint32_t* weights, int32_t* actions_out, uint8_t* flags,
                    uint8_t* valid, void* stream)
{
    return step_common(h, true, step_idx, weights, nullptr, actions_out, flags, valid, stream);
}

int pkb_sample_actions(pkb_env* h, uint32_t step_idx, const uint32_t* weights, int32_t* actions_out, void* stream)
{
    if (!h || !actions_out) return fail("pkb_sample_actions: null argument");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (weights) PKB_CUDA(cudaMemcpyAsync(h->d_weights, weights, NUM_MOVES * 4, cudaMemcpyHostToDevice, s));
    PKB_DISPATCH_P(h->cfg.num_players, (sample_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(
                                           h->d, h->gc, step_idx, weights ? 1u : 0u, h->d_weights, actions_out)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_valid_actions(pkb_env* h, int32_t player, uint8_t* valid, void* stream)
{
    if (!h || !valid) return fail("pkb_valid_actions: null argument");
    if (player >= h->cfg.num_players) return fail("pkb_valid_actions: player out of range");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_DISPATCH_P(h->cfg.num_players, (valid_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, player, valid)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_game_over(pkb_env* h, uint8_t* out, void* stream)
{
    if (!h || !out) return fail("pkb_game_over: null argument");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_DISPATCH_P(h->cfg.num_players, (game_over_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, out)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_state(pkb_env* h, pkb_state_view* out)
{
    if (!h || !out) return fail("pkb_state: null argument");
    out->num_envs = h->d.n;
    out->stride = h->d.stride;
    out->num_players = h->cfg.num_players;
    out->reserved = 0;
    out->credits = h->d.cr;
    out->bets = h->d.bt;
    out->pending_bets = h->d.pd;
    out->payoffs = h->d.payoffs;
    out->minimum_raise_value = h->d.mr;
    out->player_states = h->d.st;
    out->meta = h->d.meta;
    out->hand = h->d.hand;
    out->episode = h->d.episode;
    return 0;
}

int pkb_cards(pkb_env* h, uint8_t* out, void* stream)
{
    if (!h || !out) return fail("pkb_cards: null argument");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_DISPATCH_P(h->cfg.num_players, (cards_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, h->gc, out)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_hand_features(pkb_env* h, int32_t player, uint32_t* out, void* stream)
{
    if (!h || !out) return fail("pkb_hand_features: null argument");
    if (player >= h->cfg.num_players) return fail("pkb_hand_features: player out of range");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_DISPATCH_P(h->cfg.num_players,
                   (features_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, h->gc, player, out)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_stats(pkb_env* h, pkb_stats_t* out, void* stream)
{
    if (!h || !out) return fail("pkb_stats: null argument");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned long long c[STAT_COUNT];
    double ps[MAX_PLAYERS];
    PKB_CUDA(cudaMemcpyAsync(c, h->d.stats, sizeof c, cudaMemcpyDeviceToHost, s));
    PKB_CUDA(cudaMemcpyAsync(ps, h->d.payoff_sum, sizeof ps, cudaMemcpyDeviceToHost, s));
    PKB_CUDA(cudaStreamSynchronize(s));
    out->steps = c[STAT_STEPS];
    out->hands = c[STAT_HANDS];
    out->showdowns = c[STAT_SHOWDOWNS];
    out->games = c[STAT_GAMES];
    out->invalid = c[STAT_INVALID];
    out->poisoned = c[STAT_POISONED];
    for (int p = 0; p < PKB_MAX_PLAYERS; ++p) out->payoff_sum[p] = ps[p];
    return 0;
}

int pkb_stats_reset(pkb_env* h, void* stream)
{
    if (!h) return fail("pkb_stats_reset: null handle");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    PKB_CUDA(cudaMemsetAsync(h->d.stats, 0, STAT_COUNT * 8 + MAX_PLAYERS * 8, static_cast<cudaStream_t>(stream)));
    return 0;
}

int pkb_deal(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t hand, int32_t num_players, uint8_t* out)
{
    if (!out || num_players < 1 || 5 + 2 * num_players > 52) return fail("pkb_deal: bad argument");
    Dealer d(seed, env_id, episode, hand);
    for (int i = 0; i < 5 + 2 * num_players; ++i) out[i] = (uint8_t)card_value_from_id(d.next());
    return 0;
}

// ---- judger ----------------------------------------------------------------------------------

static unsigned eval_grid(uint64_t n, unsigned threads)
{
    const uint64_t want = (n + threads - 1) / threads;
    const uint64_t cap = 148ull * 32;  // grid-stride beyond a few waves of the 148 SMs
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

int pkb_eval_cards(const uint8_t* cards, int32_t k, uint64_t n, uint32_t* out, void* stream)
{
    if (k < 0 || k > 7) return fail("pkb_eval_cards: k must be in 0..7");
    if (n == 0) return 0;
    if (!out || (k > 0 && !cards)) return fail("pkb_eval_cards: null argument");
    eval_cards_kernel<<<eval_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cards, k, n, out);
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_eval_masks(const uint64_t* masks, uint64_t n, uint32_t* out, void* stream)
{
    if (n == 0) return 0;
    if (!out || !masks) return fail("pkb_eval_masks: null argument");
    eval_masks_kernel<<<eval_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(masks, n, out);
    PKB_CUDA(cudaGetLastError());
    return 0;
}

static cudaError_t ensure_binom()
{
    static bool done = false;  // per process; constant memory is per device context
    static int done_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (done && done_device == dev) return cudaSuccess;
    uint64_t b[53][8];
    for (int n = 0; n <= 52; ++n)
        for (int k = 0; k < 8; ++k) b[n][k] = k == 0 ? 1 : (n == 0 ? 0 : b[n - 1][k - 1] + b[n - 1][k]);
    cudaError_t e = cudaMemcpyToSymbol(c_binom, b, sizeof b);
    if (e == cudaSuccess) {
        done = true;
        done_device = dev;
    }
    return e;
}

int pkb_eval_exhaustive(int32_t k, uint64_t first, uint64_t count, uint64_t* hist, uint64_t* sums, uint32_t* out,
                        void* stream)
{
    if (k < 1 || k > 7) return fail("pkb_eval_exhaustive: k must be in 1..7");
    uint64_t total = 1;
    for (int i = 1; i <= k; ++i) total = total * (uint64_t)(52 - k + i) / (uint64_t)i;
    if (first > total || count > total - first) return fail("pkb_eval_exhaustive: range exceeds C(52,k)");
    PKB_CUDA(ensure_binom());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned long long* d_acc = nullptr;
    PKB_CUDA(cudaMalloc(&d_acc, 13 * 8));
    cudaError_t e = cudaMemsetAsync(d_acc, 0, 13 * 8, s);
    if (e == cudaSuccess && count) {
        const uint64_t threads = (count + EXH_CHUNK - 1) / EXH_CHUNK;
        eval_exhaustive_kernel<<<blocks_for(threads, 256), 256, 0, s>>>(k, first, count, out, d_acc, d_acc + 11);
        e = cudaGetLastError();
    }
    unsigned long long host[13];
    if (e == cudaSuccess) e = cudaMemcpyAsync(host, d_acc, sizeof host, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_acc);
    if (e != cudaSuccess) return fail("pkb_eval_exhaustive", e);
    if (hist) for (int i = 0; i < 11; ++i) hist[i] = host[i];
    if (sums) { sums[0] = host[11]; sums[1] = host[12]; }
    return 0;
}

int pkb_random_hands(uint64_t seed, uint64_t n, uint64_t* masks, void* stream)
{
    if (n == 0) return 0;
    if (!masks) return fail("pkb_random_hands: null argument");
    random_hands_kernel<<<eval_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, n, masks);
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_compare_rankings(const uint32_t* packed, int32_t p, uint64_t n, uint8_t* onehot, void* stream)
{
    if (p < 1 || p > 32) return fail("pkb_compare_rankings: p must be in 1..32");
    if (n == 0) return 0;
    if (!packed || !onehot) return fail("pkb_compare_rankings: null argument");
    compare_kernel<<<eval_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(packed, p, n, onehot);
    PKB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
