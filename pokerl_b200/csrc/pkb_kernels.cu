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
#include <cstdlib>
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
    uint32_t* ctl;              // [stride] automaton control word of envs parked between the step kernels
    uint32_t* list_e;           // [stride] envs whose hand ended in this step (work list of kernel B)
    uint32_t* list_s;           // [stride] envs that need a ranked showdown (work list of kernel C)
    uint32_t* counters;         // [4] list lengths: {e, s} x step parity
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
    uint32_t parity;  // which pair of list counters this step uses; the other pair is zeroed for the next step
    // the slice of the handle's envs this launch works on (slices run concurrently on internal streams)
    uint32_t env_lo, env_cnt;  // first local env, number of env slots (a multiple of the block size)
    uint32_t* list_e;          // [env_cnt] work list of kernel B
    uint32_t* list_s;          // [env_cnt] work list of kernel C
    uint32_t* counters;        // [4] list lengths: {e, s} x step parity
};

constexpr int TILE = 256;         // env-count granularity of the HBM planes (stride is a multiple of it)
constexpr int STEP_THREADS = 128; // block size of the three step kernels (one env per thread)
constexpr int MAX_SLICES = 16;    // env slices of one handle that may be stepped concurrently
#ifndef PKB_BC_MIN_BLOCKS
#define PKB_BC_MIN_BLOCKS 4
#endif
constexpr int BC_MIN_BLOCKS = PKB_BC_MIN_BLOCKS;  // resident blocks/SM kernels B and C are compiled for (register cap)

// bytes of shared memory per thread-slot: f64 planes (3P+1), meta/hand u32, P state bytes
template <int P>
__host__ __device__ constexpr size_t slot_bytes()
{
    return (size_t)(3 * P + 1) * 8 + 2 * 4 + P;
}

PKB_D View<0> global_view(const DevState& d)
{
    return View<0>{d.cr, d.bt, d.pd, d.mr, d.st, d.meta, d.hand, (size_t)d.stride};
}

// Per-thread staging of one env in shared memory ([field][thread] layout: bank-conflict free, and
// seats can be indexed dynamically without spilling register arrays to local memory)
template <int P>
PKB_D View<STEP_THREADS> carve_view(unsigned char* smem)
{
    View<STEP_THREADS> v;
    double* f64 = reinterpret_cast<double*>(smem);
    v.cr = f64;
    v.bt = f64 + P * STEP_THREADS;
    v.pd = f64 + 2 * P * STEP_THREADS;
    v.mr = f64 + 3 * P * STEP_THREADS;
    uint32_t* u32 = reinterpret_cast<uint32_t*>(f64 + (3 * P + 1) * STEP_THREADS);
    v.meta = u32;
    v.hand = u32 + STEP_THREADS;
    v.st = reinterpret_cast<uint8_t*>(u32 + 2 * STEP_THREADS);
    v.ld_rt = STEP_THREADS;
    return v;
}

// HBM planes -> the thread's staging slot; the seat states come back as masks (registers)
template <int P>
PKB_D Seats load_env(View<STEP_THREADS>& v, int t, const DevState& d, size_t env)
{
    const View<0> g = global_view(d);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        v.cr[p * STEP_THREADS + t] = d.cr[p * d.stride + env];
        v.bt[p * STEP_THREADS + t] = d.bt[p * d.stride + env];
        v.pd[p * STEP_THREADS + t] = d.pd[p * d.stride + env];
    }
    v.mr[t] = d.mr[env];
    v.meta[t] = d.meta[env];
    v.hand[t] = d.hand[env];
    return load_seats<P>(g, env);
}

template <int P>
PKB_D void store_env(const View<STEP_THREADS>& v, int t, const Seats& x, const DevState& d, size_t env)
{
    View<0> g = global_view(d);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        d.cr[p * d.stride + env] = v.cr[p * STEP_THREADS + t];
        d.bt[p * d.stride + env] = v.bt[p * STEP_THREADS + t];
        d.pd[p * d.stride + env] = v.pd[p * STEP_THREADS + t];
    }
    d.mr[env] = v.mr[t];
    d.meta[env] = v.meta[t];
    d.hand[env] = v.hand[t];
    store_seats<P>(g, env, x);
}

// block-aggregated append of env ids to a work list: one global atomic per block, entries in
// thread order.  Must be reached by every thread of the block.
PKB_D void list_append(uint32_t* list, uint32_t* counter, bool mine, uint32_t env)
{
    __shared__ uint32_t s_warp[STEP_THREADS / 32], s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned vote = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) s_warp[wid] = (uint32_t)__popc(vote);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int i = 0; i < STEP_THREADS / 32; ++i) total += s_warp[i];
        s_base = total ? atomicAdd(counter, total) : 0u;
    }
    __syncthreads();
    if (mine) {
        uint32_t at = s_base + (uint32_t)__popc(vote & ((1u << lane) - 1u));
        for (int i = 0; i < wid; ++i) at += s_warp[i];
        list[at] = env;
    }
    __syncthreads();  // s_warp / s_base are reused by the next call
}

// warp-aggregated event counter -> one global atomic per warp that saw the event
PKB_D void count_events(unsigned long long* counter, bool event)
{
    const unsigned vote = __ballot_sync(0xffffffffu, event);
    if (vote && (threadIdx.x & 31) == 0) atomicAdd(counter, (unsigned long long)__popc(vote));
}

// Step outputs of one finished env: flag bits and the valid mask of the new active player
template <int P>
PKB_D uint32_t write_outputs(const View<STEP_THREADS>& v, int t, const Ctl& c, const StepArgs& a, size_t env)
{
    const uint32_t f = (uint32_t)c.done | (c.invalid ? F_INVALID : 0u) | (c.poisoned ? F_POISONED : 0u);
    if (a.flags) a.flags[env] = (uint8_t)f;
    if (a.valid) a.valid[env] = (uint8_t)valid_mask<P>(v, t, unpack_meta(v.meta[t]).act);
    return f;
}

// What follows Game.end_hand once the payoffs are in the pending plane (game.py:530-539 and the
// caller's use of the return value), then Game.reset() if the game ended and auto-reset is on.
// Leaves c.pc == PC_END_HAND when ANOTHER hand ends inside the same step (rare: all-in blinds).
template <int P>
PKB_D void finish_hand(View<STEP_THREADS>& v, int t, Seats& x, Meta& m, Ctl& c, const StepArgs& a, size_t env,
                       double (&paysum)[P])
{
    const DevState& d = a.d;
    double pay[P];
    end_hand_finish<P>(v, t, x, m, c, d.payoffs, d.stride, env, pay, a.gc);
#pragma unroll
    for (int p = 0; p < P; ++p) paysum[p] += pay[p];
    if (c.pc == PC_DONE && (c.done & (int)F_GAME_OVER) && !c.poisoned && a.gc.auto_reset) {
        d.episode[env] += 1;  // Game.reset() right after the step that ended the game (game.py:397-412)
        reset_env<P>(v, t, x, m, a.gc);
    }
}

// game.py:473 `assert num_potential_winners > 0` would fire: payoffs were zeroed at :468; the step
// reports hand-over (and round-over when end_hand was entered from next_turn) plus PKB_F_POISONED.
template <int P>
PKB_D void assert_exit(const Seats& x, Ctl& c, const StepArgs& a, size_t env)
{
#pragma unroll
    for (int p = 0; p < P; ++p) a.d.payoffs[p * a.d.stride + env] = 0.0;
    c.done = c.cont ? (int)((game_over<P>(x) ? F_GAME_OVER : 0u) | F_HAND_OVER | F_ROUND_OVER) : (int)F_HAND_OVER;
    c.pc = PC_DONE;
}

// block reduction of the per-thread payoff sums and counters into the handle's statistics
template <int P>
PKB_D void reduce_stats(const DevState& d, const double (&paysum)[P], uint32_t hands, uint32_t shows, uint32_t games,
                        uint32_t poisoned)
{
    __shared__ double s_pay[MAX_PLAYERS];
    __shared__ uint32_t s_cnt[4];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < MAX_PLAYERS) s_pay[tid] = 0.0;
    if (tid < 4) s_cnt[tid] = 0;
    __syncthreads();
    const uint32_t cnt[4] = {hands, shows, games, poisoned};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, cnt[i]);
        if (lane == 0 && w) atomicAdd(&s_cnt[i], w);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        double x = paysum[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x != 0.0) atomicAdd(&s_pay[p], x);
    }
    __syncthreads();
    if (tid == 0 && s_cnt[0]) atomicAdd(&d.stats[STAT_HANDS], (unsigned long long)s_cnt[0]);
    if (tid == 1 && s_cnt[1]) atomicAdd(&d.stats[STAT_SHOWDOWNS], (unsigned long long)s_cnt[1]);
    if (tid == 2 && s_cnt[2]) atomicAdd(&d.stats[STAT_GAMES], (unsigned long long)s_cnt[2]);
    if (tid == 3 && s_cnt[3]) atomicAdd(&d.stats[STAT_POISONED], (unsigned long long)s_cnt[3]);
    if (tid >= 32 && tid < 32 + P && s_pay[tid - 32] != 0.0) atomicAdd(&d.payoff_sum[tid - 32], s_pay[tid - 32]);
}

// --------------------------------------------------------------------------------------------
// Game.step (game.py:621-700) as a three-kernel pipeline over the SoA planes.  The branches of a
// step differ by 10-100x in cost (a bet; a hand that ends; a ranked showdown), so instead of letting
// warps diverge, each kernel runs ONE kind of work at full warp width and hands the rest on
// through a compacted env list in HBM:
//   A  every env:          validate/sample the action, bet, seat scan, next_turn commits  -> list E
//   B  envs in list E:     end_hand: commit, last-stand pots, deal the next hand, resume    -> list S
//   C  envs in list S:     ranked showdown (deal replay, 7-card ranks, side pots), then as B
// Each kernel is small (instruction-cache resident) and streams state through shared memory.
// --------------------------------------------------------------------------------------------

// K1: one thread per env, fully coalesced loads/stores
template <int P, bool SAMPLE>
__global__ void __launch_bounds__(STEP_THREADS) step_kernel_a(const __grid_constant__ StepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    View<STEP_THREADS> v = carve_view<P>(smem_raw);
    const DevState& d = a.d;
    const int t = threadIdx.x;
    const size_t env = (size_t)a.env_lo + (size_t)blockIdx.x * STEP_THREADS + t;
    const bool live = env < d.n;
    if (blockIdx.x == 0 && t < 2) a.counters[2 * (a.parity ^ 1u) + t] = 0;  // next step's list lengths

    Seats x = load_env<P>(v, t, d, env);  // stride is a multiple of the block size: always in bounds
    Ctl c{PC_DONE, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
        Meta m = unpack_meta(v.meta[t]);
        const uint32_t vm = valid_mask<P>(v, t, m.act);
        int action;
        if (SAMPLE) {
            action = sample_action(a.gc.seed, a.gc.env_id_base + (uint32_t)env, a.step_idx, vm,
                                   a.use_weights ? a.weights : nullptr);
            if (a.actions_out) a.actions_out[env] = action;
        } else
            action = a.actions[env];
        step_begin<P>(v, t, x, m, c, action, vm, a.gc);
        v.meta[t] = pack_meta(m);
    }
    const bool parked = live && c.pc == PC_END_HAND;
    if (live) {
        if (!c.invalid) store_env<P>(v, t, x, d, env);  // an invalid action changes nothing (game.py:649-651)
        if (parked) d.ctl[env] = pack_ctl(c);
        else write_outputs<P>(v, t, c, a, env);
    }
    list_append(a.list_e, &a.counters[2 * a.parity], parked, (uint32_t)env);
    if (a.gc.track_stats) {
        count_events(&d.stats[STAT_INVALID], live && c.invalid);
        count_events(&d.stats[STAT_POISONED], live && c.poisoned);
        if (blockIdx.x == 0 && t == 0) {  // live envs of this slice; pkb_stats subtracts the invalid ones
            const size_t hi = (size_t)a.env_lo + a.env_cnt;
            atomicAdd(&d.stats[STAT_STEPS], (unsigned long long)((hi < d.n ? hi : d.n) - a.env_lo));
        }
    }
}

// K2a: end_hand for the envs in list E
template <int P>
__global__ void __launch_bounds__(STEP_THREADS, BC_MIN_BLOCKS) step_kernel_b(const __grid_constant__ StepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    View<STEP_THREADS> v = carve_view<P>(smem_raw);
    const DevState& d = a.d;
    const int t = threadIdx.x;
    const uint32_t n_items = a.counters[2 * a.parity];
    uint32_t games = 0, poisoned = 0;
    double paysum[P];
#pragma unroll
    for (int p = 0; p < P; ++p) paysum[p] = 0.0;
    const uint32_t rounded = (n_items + STEP_THREADS - 1) / STEP_THREADS * STEP_THREADS;
    for (uint32_t i = blockIdx.x * STEP_THREADS + t; i < rounded; i += gridDim.x * STEP_THREADS) {
        const bool active = i < n_items;
        bool to_c = false;
        size_t env = 0;
        if (active) {
            env = a.list_e[i];
            Seats x = load_env<P>(v, t, d, env);
            Ctl c = unpack_ctl(d.ctl[env]);
            Meta m = unpack_meta(v.meta[t]);
            if (end_hand_begin<P>(v, t, x, c)) {
                c.pc = PC_SHOWDOWN;
                to_c = true;
            } else if (c.poisoned) {
                assert_exit<P>(x, c, a, env);
            } else {
                finish_hand<P>(v, t, x, m, c, a, env, paysum);
                to_c = c.pc == PC_END_HAND;  // another hand ends inside this step: kernel C runs it
            }
            v.meta[t] = pack_meta(m);
            store_env<P>(v, t, x, d, env);
            if (to_c) d.ctl[env] = pack_ctl(c);
            else {
                const uint32_t f = write_outputs<P>(v, t, c, a, env);
                games += (f & F_GAME_OVER) ? 1u : 0u;
                poisoned += c.poisoned ? 1u : 0u;
            }
        }
        list_append(a.list_s, &a.counters[2 * a.parity + 1], to_c, (uint32_t)env);
    }
    if (a.gc.track_stats) {
        const uint32_t hands = (blockIdx.x == 0 && t == 0) ? n_items : 0u;
        reduce_stats<P>(d, paysum, hands, 0u, games, poisoned);
    }
}

// K2b: ranked showdowns (and the rare further hands of a step) for the envs in list S
template <int P>
__global__ void __launch_bounds__(STEP_THREADS, BC_MIN_BLOCKS) step_kernel_c(const __grid_constant__ StepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    View<STEP_THREADS> v = carve_view<P>(smem_raw);
    const DevState& d = a.d;
    const int t = threadIdx.x;
    const uint32_t n_items = a.counters[2 * a.parity + 1];
    uint32_t hands = 0, shows = 0, games = 0, poisoned = 0;
    double paysum[P];
#pragma unroll
    for (int p = 0; p < P; ++p) paysum[p] = 0.0;
    for (uint32_t i = blockIdx.x * STEP_THREADS + t; i < n_items; i += gridDim.x * STEP_THREADS) {
        const size_t env = a.list_s[i];
        Seats x = load_env<P>(v, t, d, env);
        Ctl c = unpack_ctl(d.ctl[env]);
        Meta m = unpack_meta(v.meta[t]);
        const uint32_t episode = d.episode[env];
        bool asserted = false;
        while (c.pc != PC_DONE) {
            bool ranked = c.pc == PC_SHOWDOWN;
            if (!ranked) {  // PC_END_HAND: a further hand of this step, from its beginning
                ++hands;
                ranked = end_hand_begin<P>(v, t, x, c);
                if (!ranked && c.poisoned) {
                    asserted = true;
                    break;
                }
            }
            if (ranked) {
                ++shows;
                showdown<P>(v, t, x, c, a.gc.seed, a.gc.env_id_base + (uint32_t)env, episode,
                            [](uint64_t mask) { return eval_mask52(mask); });
            }
            c.pc = PC_DONE;
            finish_hand<P>(v, t, x, m, c, a, env, paysum);
        }
        if (asserted) assert_exit<P>(x, c, a, env);
        v.meta[t] = pack_meta(m);
        store_env<P>(v, t, x, d, env);
        const uint32_t f = write_outputs<P>(v, t, c, a, env);
        games += (f & F_GAME_OVER) ? 1u : 0u;
        poisoned += c.poisoned ? 1u : 0u;
    }
    if (a.gc.track_stats) reduce_stats<P>(d, paysum, hands, shows, games, poisoned);
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
    Seats x{0, 0, 0, 0};
    reset_env<P>(v, env, x, m, gc);
    store_seats<P>(v, env, x);
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
    out[env] = game_over<P>(load_seats<P>(global_view(d), env)) ? 1 : 0;
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
    uint32_t parity;  // step parity: which pair of list counters the next step uses
    // env slices stepped concurrently on internal streams (fork/join around the caller's stream)
    int num_slices;
    cudaStream_t streams[MAX_SLICES];
    cudaEvent_t fork, join[MAX_SLICES];
};

static thread_local std::string g_error;

static int slices_from_env()
{
    const char* e = getenv("PKB_SLICES");
    int n = e ? atoi(e) : 1;
    return n < 1 ? 1 : (n > MAX_SLICES ? MAX_SLICES : n);
}

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
    constexpr size_t smem = slot_bytes<P>() * STEP_THREADS;
    static int sm_count[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sm_count[dev]) {
        e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    const unsigned blocks_a = a.env_cnt / STEP_THREADS;
    // B and C read their item count on the device.  Their grids cover the usual share of envs whose
    // hand ends (about 1/3) / is ranked (about 1/8) in one pass, so that all items are in flight at
    // once; a grid stride handles anything beyond that, and surplus blocks exit at once.
    const unsigned blocks_b = blocks_a / 2 + 1, blocks_c = blocks_a / 4 + 1;
    step_kernel_a<P, SAMPLE><<<blocks_a, STEP_THREADS, smem, s>>>(a);
    step_kernel_b<P><<<blocks_b, STEP_THREADS, smem, s>>>(a);
    step_kernel_c<P><<<blocks_c, STEP_THREADS, smem, s>>>(a);
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
    const size_t n_f64 = (size_t)(4 * P + 1) * stride, n_u32 = (size_t)6 * stride, n_u8 = (size_t)P * stride;
    const size_t off_u32 = n_f64 * 8, off_u8 = off_u32 + n_u32 * 4, off_stats = (off_u8 + n_u8 + 255) / 256 * 256;
    h->arena_bytes = off_stats + STAT_COUNT * 8 + MAX_PLAYERS * 8 + NUM_MOVES * 4 + 4 + 4 * 4 * MAX_SLICES;
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
    d.ctl = u + 3 * stride;
    d.list_e = u + 4 * stride;
    d.list_s = u + 5 * stride;
    d.st = reinterpret_cast<uint8_t*>(base + off_u8);
    d.stats = reinterpret_cast<unsigned long long*>(base + off_stats);
    d.payoff_sum = reinterpret_cast<double*>(base + off_stats + STAT_COUNT * 8);
    h->d_weights = reinterpret_cast<uint32_t*>(base + off_stats + STAT_COUNT * 8 + MAX_PLAYERS * 8);
    d.counters = h->d_weights + NUM_MOVES + 1;
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
    // slices: at least 64 blocks each, so small batches stay on the caller's stream
    h->num_slices = slices_from_env();
    while (h->num_slices > 1 && stride / h->num_slices < 64 * STEP_THREADS) --h->num_slices;
    if (h->num_slices > 1) {
        e = cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming);
        for (int i = 0; i < h->num_slices && e == cudaSuccess; ++i) {
            e = cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->join[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) {
            cudaFree(h->arena);
            delete h;
            return fail("pkb_create: streams", e);
        }
    }
    *out = h;
    return 0;
}

int pkb_destroy(pkb_env* h)
{
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    if (h->num_slices > 1) {
        for (int i = 0; i < h->num_slices; ++i) {
            cudaStreamSynchronize(h->streams[i]);
            cudaStreamDestroy(h->streams[i]);
            cudaEventDestroy(h->join[i]);
        }
        cudaEventDestroy(h->fork);
    }
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
    a.parity = h->parity;
    h->parity ^= 1u;
    for (int i = 0; i < NUM_MOVES; ++i) a.weights[i] = weights ? weights[i] : 1u;
    cudaStream_t caller = static_cast<cudaStream_t>(stream);
    const int S = h->num_slices;
    // slice boundaries on multiples of the block size
    const uint64_t blocks = h->d.stride / STEP_THREADS, per = (blocks + S - 1) / S;
    if (S > 1) PKB_CUDA(cudaEventRecord(h->fork, caller));
    for (int i = 0; i < S; ++i) {
        const uint64_t b0 = per * i, b1 = b0 + per < blocks ? b0 + per : blocks;
        if (b0 >= b1) continue;
        a.env_lo = (uint32_t)(b0 * STEP_THREADS);
        a.env_cnt = (uint32_t)((b1 - b0) * STEP_THREADS);
        a.list_e = h->d.list_e + a.env_lo;
        a.list_s = h->d.list_s + a.env_lo;
        a.counters = h->d.counters + 4 * i;
        cudaStream_t s = S > 1 ? h->streams[i] : caller;
        if (S > 1) PKB_CUDA(cudaStreamWaitEvent(s, h->fork, 0));
        cudaError_t e = cudaSuccess;
        PKB_DISPATCH_P(h->cfg.num_players,
                       e = sample ? launch_step<P, true>(a, s) : launch_step<P, false>(a, s));
        if (e != cudaSuccess) return fail("pkb_step: launch", e);
        if (S > 1) {
            PKB_CUDA(cudaEventRecord(h->join[i], s));
            PKB_CUDA(cudaStreamWaitEvent(caller, h->join[i], 0));
        }
    }
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
    out->steps = c[STAT_STEPS] - c[STAT_INVALID];  // kernel A adds num_envs per launch
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
