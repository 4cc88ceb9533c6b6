// pkb_kernels.cu -- CUDA kernels (sm_100a) and the C ABI of include/pokerl_b200.h.
//
// Kernels
//   step_warp_kernel K1+K2(+K5,K6) Game.step: a 128-env tile per warp held in shared memory (2-D TMA tensor copies of
//                                the SoA plane rows), passes over warp-compacted slot lists (every env /
//                                showdown list / end_hand list)
//   reset_kernel     K3          Game.reset / setup_hand (dealing is a pure Philox function)
//   eval_*_kernel    K4          eval_hand on card lists / id masks / all C(52,k) hands
//   sample_kernel    K5          RandomAgent over the valid mask
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/pokerl_b200.h"
#include "pkb_game.cuh"

namespace pkb {

// --------------------------------------------------------------------------------------------
// device-side state (one arena per handle)
// --------------------------------------------------------------------------------------------
struct DevState {
    double *cr, *bt, *pd, *mr, *payoffs;
    uint64_t* sm;  // seat masks (player_states), one 64-bit word per env: the row after mr
    uint32_t *meta, *hand, *episode;
    unsigned long long* stats;  // [STAT_COUNT] counters
    double* payoff_sum;         // [MAX_PLAYERS]
    // Cross-launch pipelining of the step kernel (programmatic dependent launch): a step launch may start while
    // the previous one is still draining, so a tile is guarded by its own epoch word instead of by kernel order.
    unsigned long long* seq;    // [1] blocks of step launches started so far: block ticket / grid size = the launch's index
    uint32_t* epoch;            // [tiles] number of step launches that have stored this tile (release/acquire)
    uint64_t n, stride;
};

enum : int { STAT_STEPS = 0, STAT_HANDS, STAT_SHOWDOWNS, STAT_GAMES, STAT_INVALID, STAT_POISONED, STAT_COUNT };

// TMA descriptors of the two plane groups of a handle (2-D: [rows][stride]), box = [rows][warp tile]
struct alignas(64) TileMaps {
    CUtensorMap w64;  // credits, bets, pending_bets ([P] rows each), minimum_raise_value, seat masks: 3P + 2 rows of 8-byte words
    CUtensorMap u32;  // meta, hand, episode
};

struct StepArgs {
    TileMaps maps;
    DevState d;
    GameConst gc;
    const int32_t* actions;
    const uint8_t* actions_u8;  // compact form of `actions` (255 = skip); used when `actions` is null
    const uint8_t* actions_u4;  // two moves per byte (env 2i in the low nibble; 15 = skip); used when both others are null
    uint8_t* packed;            // or null: one output byte per env, flags bits 0-2 | valid bits of moves 1..5 << 3
    int32_t* actions_out;
    uint8_t* flags;
    uint8_t* valid;
    uint32_t step_idx;
    uint32_t reps;       // device-sampled steps per tile visit (pkb_step_random_n): the tile stays in shared memory between them
    uint32_t use_weights;
    uint32_t weights[NUM_MOVES];
    uint32_t num_tiles;
    uint32_t wait_prev;  // the launch before this one on the stream is not a step launch of this handle: wait for it
    // PokerGameEnv mode (step_warp_kernel<P, false, true>): see env_advance
    uint32_t env_reset;         // 1: PokerGameEnv.reset for the envs selected by env_mask; 0: PokerGameEnv.step
    uint32_t round_base;        // the j-th opponent move of an env inside this call draws with step index round_base + j - 1
    const uint8_t* env_mask;    // reset: envs to reset (null: all)
    double* reward;             // step outputs, one per env
    uint8_t* env_done;          //   bit 0 done, bit 1 hand_over, bit 5 not stepped (PKB_ACTION_SKIP), bit 6 rejected, bit 7 poisoned
    // PokerGameEnv.step in phases (env_step_phases): a launch plays at most env_round_cap opponent rounds and hands the envs
    // whose loop has not returned to the next phase, which runs them packed densely in a scratch arena
    uint32_t env_round_cap;      // opponent rounds this launch may play (0: no limit but ENV_ROUND_CAP)
    uint32_t env_round_offset;   // continue mode: opponent moves every env of this launch has already made in this call
    uint32_t no_epoch;           // scratch arena: plain stream order, no tile hand-over between launches
    const uint32_t* env_ids;     // continue mode: slot i of this launch is env env_ids[i] of the handle (null: identity)
    const uint32_t* n_dev;       // continue mode: the number of envs of this launch, on the device
    const uint8_t* env_ctl_in;   // continue mode: the control byte of every slot as the previous phase left it
    uint32_t* strag_count;       // out (null: none): envs still pending when the cap is reached -- how many,
    uint32_t* strag_ids;         //   which (indices into the handle's planes)
    uint8_t* strag_ctl;          //   and their control bytes
};

#ifndef PKB_STRIDE_ALIGN
#define PKB_STRIDE_ALIGN 1024
#endif
constexpr int STRIDE_ALIGN = PKB_STRIDE_ALIGN;  // env-count granularity of the HBM planes: a multiple of every tile size

PKB_D View<0> global_view(const DevState& d)
{
    return View<0>{d.cr, d.bt, d.pd, d.mr, d.sm, d.meta, d.hand, (size_t)d.stride};
}

PKB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

PKB_D void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PKB_D void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifndef PKB_MBAR_HINT_NS
#define PKB_MBAR_HINT_NS 20000
#endif
PKB_D bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"((uint32_t)PKB_MBAR_HINT_NS)  // suspend-time hint (ns): do not spin on issue slots
        : "memory");
    return ok != 0;
}
// Watchdog: a tile load that has not landed after 2^18 polls (milliseconds at least) never will; abort the
// launch with an error instead of hanging the device.
PKB_D void mbar_wait(uint32_t bar, uint32_t parity)
{
#ifdef PKB_SPIN_WAIT  // measurement: poll without suspending
    uint32_t ok = 0;
    for (uint32_t polls = 0; !ok; ++polls) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (polls > (1u << 26)) __trap();
    }
#else
    for (uint32_t polls = 0; !mbar_try_wait(bar, parity); ++polls)
        if (polls > (1u << 18)) __trap();
#endif
}
// TMA 2-D tiled copies (SASS: UTMALDG / UTMASTG): box [rows][tile envs] <-> dense shared-memory rows
PKB_D void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
PKB_D void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src_smem)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(src_smem)
                 : "memory");
}
PKB_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
PKB_D void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
PKB_D void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
PKB_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
PKB_D void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// programmatic dependent launch: lets the next kernel of the stream start once every block of this one has got here
PKB_D void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
PKB_D uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
PKB_D void st_release_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// Tile hand-over between two step launches that are in flight at the same time.  The launch with index L may load a
// tile once L launches have stored it; it then publishes L + 1.  The producer is a block of an EARLIER launch of the
// same stream, and every block of an earlier launch is resident before a later launch starts (that is the launch
// condition of griddepcontrol), so the wait cannot deadlock; the poll cap turns a broken chain (a failed launch,
// two streams racing on one handle) into a launch error instead of a hang.
PKB_D void epoch_wait(const uint32_t* word, uint32_t want)
{
    for (uint32_t polls = 0; ld_acquire_u32(word) != want; ++polls) {
        __nanosleep(64);
        if (polls > (1u << 22)) __trap();
    }
}
// The flag store must come after (1) the tile's rows, written by this thread's TMA stores, and (2) the warp's plain
// stores to the caller's output arrays and the payoff plane, which the next launch may overwrite.  `wait_group 0` alone
// is NOT enough for (1): it returns when the writes are performed with respect to this thread, and a consumer on
// another SM that sees a relaxed flag store a moment later can still load the old rows (measured: a 2^18-env batch,
// whose launches overlap three deep, diverges from the oracle within 48 steps; tests/test_gpu_round2.py).  So the flag
// is a release store at GPU scope (MEMBAR.ALL.GPU in SASS, about 1.5 us of a 38 us step).
PKB_D void epoch_publish(uint32_t* word, uint32_t value)
{
    bulk_wait_all();
    fence_async_global();
    st_release_u32(word, value);
}

template <class L>
PKB_D View<L::T> tile_view(unsigned char* smem)
{
    View<L::T> v;
    v.cr = reinterpret_cast<double*>(smem + L::OFF_CR);
    v.bt = reinterpret_cast<double*>(smem + L::OFF_BT);
    v.pd = reinterpret_cast<double*>(smem + L::OFF_PD);
    v.mr = reinterpret_cast<double*>(smem + L::OFF_MR);
    v.meta = reinterpret_cast<uint32_t*>(smem + L::OFF_META);
    v.hand = reinterpret_cast<uint32_t*>(smem + L::OFF_HAND);
    v.sm = reinterpret_cast<uint64_t*>(smem + L::OFF_SM);
    v.ld_rt = L::T;
    return v;
}

// per-thread statistics of one launch, reduced once per block
template <int P>
struct Tally {
    double pay[P];
    uint32_t hands, shows, games, poisoned, invalid, live;
};

// Step outputs of one finished env: flag bits and the valid mask of the new active player, into the tile's two output
// rows (they leave for the caller's arrays with the tile).  The valid mask is also cached in the meta word (m.vm; the
// caller stores the word): the next step's sampler draws from it.
template <int P, class V>
PKB_D uint32_t write_outputs(const V& v, int s, const Ctl& c, Meta& m, uint8_t* out_flags, uint8_t* out_valid, bool skipped = false)
{
    const uint32_t f = skipped ? F_SKIPPED : (uint32_t)c.done | (c.invalid ? F_INVALID : 0u) | (c.poisoned ? F_POISONED : 0u);
    out_flags[s] = (uint8_t)f;
    m.vm = valid_mask<P>(v, s, m.act);
    out_valid[s] = (uint8_t)m.vm;
    return f;
}

// block reduction of the per-thread tallies into the handle's statistics
template <int P>
PKB_D void reduce_stats(const DevState& d, const Tally<P>& tl)
{
    __shared__ double s_pay[MAX_PLAYERS];
    __shared__ uint32_t s_cnt[6];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < MAX_PLAYERS) s_pay[tid] = 0.0;
    if (tid < 6) s_cnt[tid] = 0;
    __syncthreads();
    const uint32_t cnt[6] = {tl.live, tl.hands, tl.shows, tl.games, tl.invalid, tl.poisoned};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, cnt[i]);
        if (lane == 0 && w) atomicAdd(&s_cnt[i], w);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        double x = tl.pay[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x != 0.0) atomicAdd(&s_pay[p], x);
    }
    __syncthreads();
    // STAT_STEPS counts every live env of the launch; pkb_stats subtracts the invalid ones
    if (tid < 6 && s_cnt[tid]) atomicAdd(&d.stats[tid], (unsigned long long)s_cnt[tid]);
    if (tid >= 32 && tid < 32 + P && s_pay[tid - 32] != 0.0) atomicAdd(&d.payoff_sum[tid - 32], s_pay[tid - 32]);
}

#ifdef PKB_TIMING
__device__ unsigned long long g_wrp[4096][16];  // per warp of the last launch: start ns, end ns, SM id, 5 phase cycle sums, tiles, C/D rounds, longest C pass, max showdowns
#endif

// --------------------------------------------------------------------------------------------
// Game.step on every env, warp-tile form (the default): the same three passes, but a tile belongs to
// ONE WARP.  A warp pulls the plane rows of its WT consecutive envs into its own shared-memory
// buffer (TMA bulk copies on the warp's own mbarrier), runs pass A over WT/32 slots per lane, builds
// the two work lists with ballots (no atomics), runs passes C and D on them, and pushes the rows back.
// There is no block barrier anywhere in the loop, so a warp whose tile holds many showdowns (or a slot
// that plays several hands in one step) delays nobody else; the 20-odd resident warps per SM cover
// each other's load latency and dependent-instruction stalls.
// --------------------------------------------------------------------------------------------
#ifndef PKB_WT_SMALL
#define PKB_WT_SMALL 128  // envs per warp tile, P <= 3
#endif
#ifndef PKB_WT_MID
#define PKB_WT_MID 128    // P == 4
#endif
#ifndef PKB_WT_LARGE
#define PKB_WT_LARGE 128  // P >= 5: list density beats warp count (measured: 128-env tiles, 6 warps/SM > 64-env tiles, 12 warps/SM)
#endif
#ifndef PKB_WT_WARPS
#define PKB_WT_WARPS 4    // warps per block (they only share the block's statistics reduction), P <= 4
#endif
#ifndef PKB_WT_UNROLL_A
#define PKB_WT_UNROLL_A 1  // pass A stays one rolled loop: the hot code has to fit the instruction cache
#endif
#define PKB_PRAGMA_(x) _Pragma(#x)
#define PKB_UNROLL(n) PKB_PRAGMA_(unroll n)
#ifndef PKB_WT_WARPS_LARGE
#define PKB_WT_WARPS_LARGE 2  // P >= 5
#endif

static_assert(STRIDE_ALIGN % PKB_WT_SMALL == 0 && STRIDE_ALIGN % PKB_WT_MID == 0 && STRIDE_ALIGN % PKB_WT_LARGE == 0, "warp tiles must divide the plane stride");

template <int P>
__host__ __device__ constexpr int wt_envs()
{
    return P <= 3 ? PKB_WT_SMALL : P == 4 ? PKB_WT_MID : PKB_WT_LARGE;
}

template <int P>
struct WarpTile {
    static constexpr int T = wt_envs<P>();
    static constexpr int NP = P;
    static constexpr int WARPS = P <= 4 ? PKB_WT_WARPS : PKB_WT_WARPS_LARGE;
    // one buffer = two dense row groups, each the box of one TMA descriptor (same row order as the HBM arena)
    static constexpr int W64_ROWS = 3 * P + 2;  // credits[P], bets[P], pending[P], minimum_raise, seat masks
    static constexpr int U32_ROWS = 3;          // meta, hand, episode
    static constexpr size_t OFF_CR = 0;
    static constexpr size_t OFF_BT = OFF_CR + (size_t)P * T * 8;
    static constexpr size_t OFF_PD = OFF_BT + (size_t)P * T * 8;
    static constexpr size_t OFF_MR = OFF_PD + (size_t)P * T * 8;
    static constexpr size_t OFF_SM = OFF_MR + (size_t)T * 8;
    static constexpr size_t OFF_META = OFF_SM + (size_t)T * 8;
    static constexpr size_t OFF_HAND = OFF_META + (size_t)T * 4;
    static constexpr size_t OFF_EP = OFF_HAND + (size_t)T * 4;
    static constexpr size_t BUF_BYTES = OFF_EP + (size_t)T * 4;  // = bytes of the two boxes
    // per warp: buffer | list E (u32: slot | control word << 8) | list S (u8 slot); the warps' mbarriers sit in the block header
    static constexpr size_t OFF_LIST_E = (BUF_BYTES + 15) / 16 * 16;
    static constexpr size_t OFF_LIST_S = OFF_LIST_E + (size_t)T * 4;
    static constexpr size_t OFF_ACT = OFF_LIST_S + (size_t)T;  // u8 per slot: the caller's actions (pkb_step), fetched while the tile loads
    static constexpr size_t OFF_OUT_F = OFF_ACT + (size_t)T;    // u8 per slot: the step's flag byte ...
    static constexpr size_t OFF_OUT_V = OFF_OUT_F + (size_t)T;  // ... and valid mask, written to the caller's arrays with the tile
    static constexpr size_t OFF_PAIRS = OFF_OUT_V + (size_t)T;  // u16 (slot | seat << 8): the showing seats of up to 32 hands (P > 2)
    static constexpr size_t OFF_PEND = OFF_PAIRS + (P > 2 ? (size_t)64 * P : 0);  // u8 per slot: the env mode's list of slots whose loop has not returned
    static constexpr size_t WARP_BYTES = (OFF_PEND + (size_t)T + 127) / 128 * 128;
    static constexpr size_t HEADER = 128;
    static constexpr size_t BYTES = HEADER + WARP_BYTES * WARPS;
    static constexpr int THREADS = 32 * WARPS;
    // resident blocks per SM that 227 KB of shared memory (1 KB reserved per block) and 64 K registers at >= 64 per thread allow
    static constexpr int BY_SMEM = (int)((227 * 1024) / (BYTES + 1024));
    static constexpr int BY_REGS = 1024 / THREADS;
    static constexpr int FIT_BLOCKS = BY_SMEM < BY_REGS ? (BY_SMEM < 1 ? 1 : BY_SMEM) : BY_REGS;
#ifdef PKB_WT_BLOCKS  // tuning: cap the resident blocks per SM (more registers per thread, fewer warps)
    static constexpr int MIN_BLOCKS = FIT_BLOCKS < PKB_WT_BLOCKS ? FIT_BLOCKS : PKB_WT_BLOCKS;
#else
    static constexpr int MIN_BLOCKS = FIT_BLOCKS;
#endif
    static_assert(OFF_META % 128 == 0 && WARPS <= 16, "TMA box destinations are 128-byte aligned");
};
static_assert(PKB_WT_SMALL <= 256 && PKB_WT_MID <= 256 && PKB_WT_LARGE <= 256, "list entries hold the slot in 8 bits");

// ENV (with SAMPLE = false): PokerGameEnv.step / reset in ONE launch -- seat 0 plays the caller's action, the other seats
// are played by the device sampler, round after round inside the tile, until the reference's loops would return
// (envs/game_env.py:20-54; the per-env control is env_advance, pkb_game.cuh).
template <int P, bool SAMPLE, bool ENV = false>
__global__ void __launch_bounds__(WarpTile<P>::THREADS, WarpTile<P>::MIN_BLOCKS) step_warp_kernel(const __grid_constant__ StepArgs a)
{
    static_assert(!(SAMPLE && ENV), "the env mode takes seat 0's action from the caller");
    using L = WarpTile<P>;
    constexpr int T = L::T;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem[];
    const DevState& d = a.d;
    // the envs of this launch: the handle's, or (env mode, continue phase) as many as the previous phase left pending
    const size_t n_envs = ENV && a.n_dev ? (size_t)*a.n_dev : d.n;
    const uint32_t num_tiles = ENV && a.n_dev ? (uint32_t)((n_envs + T - 1) / T) : a.num_tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* buf = smem + L::HEADER + (size_t)warp * L::WARP_BYTES;
    const uint32_t bar = smem_u32(smem) + 8u * (uint32_t)warp, buf_s = smem_u32(buf);
    uint32_t* list_e = reinterpret_cast<uint32_t*>(buf + L::OFF_LIST_E);
    uint8_t* list_s = buf + L::OFF_LIST_S;
    uint8_t* act_row = buf + L::OFF_ACT;
    uint8_t* out_f = buf + L::OFF_OUT_F;
    uint8_t* out_v = buf + L::OFF_OUT_V;
    uint16_t* pairs = reinterpret_cast<uint16_t*>(buf + L::OFF_PAIRS);
    uint8_t* pend = buf + L::OFF_PEND;
    // whole words of the caller's output arrays can be written at once when both arrays are 4-byte aligned
    const bool out_words = ((reinterpret_cast<uintptr_t>(a.flags) | reinterpret_cast<uintptr_t>(a.valid)) & 3u) == 0;
    View<T> v = tile_view<L>(buf);
    uint32_t* ep = reinterpret_cast<uint32_t*>(buf + L::OFF_EP);
    const uint32_t lanes_below = (1u << lane) - 1u;

    // Which launch of this handle is this?  Blocks of one launch all start before any block of the next one does
    // (stream order; with programmatic dependent launch: because launch_dependents comes after the ticket), and every
    // launch of a handle has the same grid, so the block's ticket divided by the grid size is the launch index.
    __shared__ uint32_t s_launch;
    if (threadIdx.x == 0) s_launch = (uint32_t)(atomicAdd(d.seq, 1ull) / gridDim.x);
    if (lane == 0) mbar_init(bar, 1);
    __syncthreads();
    const uint32_t launch = s_launch;
    pdl_launch_dependents();  // the next step launch may start filling the SMs this one's finished blocks leave
    // Launched with programmatic stream serialization, this grid may be running before its predecessor has completed.
    // After another step launch of this handle that is the point (tiles are handed over one by one, see epoch_wait);
    // after anything else -- a reset, a kernel of the caller's that wrote the state planes -- wait for it properly.
    if (a.wait_prev) asm volatile("griddepcontrol.wait;" ::: "memory");

    Tally<P> tl;
#pragma unroll
    for (int p = 0; p < P; ++p) tl.pay[p] = 0.0;
    tl.hands = tl.shows = tl.games = tl.poisoned = tl.invalid = tl.live = 0;

#ifdef PKB_TIMING
    unsigned long long wt_acc[6] = {0, 0, 0, 0, 0, 0};  // cycles: load wait, pass A, pass C, pass D, store; tiles
    unsigned long long wt_rounds = 0, wt_maxc = 0, wt_maxns = 0;
    unsigned long long wt_lw[3] = {0, 0, 0};  // load wait of a warp's first, second and later tiles
    unsigned long long wt_gt0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(wt_gt0));
    long long wt_last = clock64();
#define PKB_WTICK(i)                                             \
    do {                                                         \
        const long long now_ = clock64();                        \
        wt_acc[i] += (unsigned long long)(now_ - wt_last);       \
        wt_last = now_;                                          \
    } while (0)
#else
#define PKB_WTICK(i)
#endif
    // A warp's tiles are its global index plus multiples of the warp count: the same tile meets the same warp in every
    // launch, and launch L + 1 reaches a tile in the order launch L finished them.
    const uint32_t warps_total = gridDim.x * L::WARPS;
    uint32_t tile = blockIdx.x * L::WARPS + warp;
    uint32_t stored = 0xffffffffu;  // lane 0: the tile whose store is in flight and whose epoch is still to be published
    bool handed_over = false;       // lane 0: the acquire of the coming tile's epoch has already succeeded (polled a pass early)
    for (uint32_t it = 0; tile < num_tiles; ++it) {
        const size_t base = (size_t)tile * T;
        if (lane == 0) {  // two TMA box copies bring the tile in, once the previous launch has stored it
            if (!handed_over && !a.no_epoch) epoch_wait(&d.epoch[tile], launch);
            fence_async_global();  // the tile was written through the async proxy and is about to be read through it
            mbar_arrive_expect_tx(bar, (uint32_t)L::BUF_BYTES);
            tma_load_2d(buf_s + (uint32_t)L::OFF_CR, &a.maps.w64, (int)base, 0, bar);
            tma_load_2d(buf_s + (uint32_t)L::OFF_META, &a.maps.u32, (int)base, 0, bar);
        }
        const uint32_t next = tile + warps_total;
        if (!SAMPLE && !(ENV && a.env_ids)) {  // the caller's actions, as bytes (254: out of range, 255: skip), fetched while the tile is in flight
#pragma unroll
            for (int k = 0; k < T / 32; ++k) {
                const size_t env = base + (size_t)(lane + 32 * k);
                uint32_t b = 254u;
                if (env < n_envs) {
                    if (ENV && a.env_reset) {  // PokerGameEnv.reset: the byte says whether this env takes part
                        b = (!a.env_mask || a.env_mask[env]) ? 1u : 0u;
                    } else if (a.actions) {
                        const int32_t x = a.actions[env];
                        b = x == ACTION_SKIP ? 255u : (x < 0 || x > 253) ? 254u : (uint32_t)x;
                    } else if (a.actions_u8) {
                        b = a.actions_u8[env];
                    } else {
                        const uint32_t nib = (uint32_t)a.actions_u4[env >> 1] >> (4u * (uint32_t)(env & 1u)) & 15u;
                        b = nib == 15u ? 255u : nib;  // 7..14 are out of range like any other number
                    }
                }
                act_row[lane + 32 * k] = (uint8_t)b;
            }
            __syncwarp();
        }
        // While the tile is in flight -- the action draws of a lane's slots: envs whose ids differ only in bits 5-6 share a Philox block (pkb_philox.cuh),
        // so the slots lane, lane+32, ... of a tile need the block of the 128-aligned id group the tile starts in and,
        // if the tile runs past that group's end, the next group's: one or two blocks per lane and tile instead of one
        // per slot (tiles of up to 160 envs span at most two groups).
        // The words are laid out in slot order here (queue[k] = the draw of slot lane + 32k); pass A takes the head of the
        // queue and shifts it, so its rolled loop selects nothing.
        constexpr bool SHARED_DRAW = SAMPLE && T <= 160;
        constexpr int NQ = SHARED_DRAW ? T / 32 : 1;
        uint32_t queue[NQ] = {0};
        uint32_t step_now = a.step_idx, rep = 0;  // pkb_step_random_n: the step index of the repetition that is running
        auto fill_queue = [&](uint32_t step) {
            if constexpr (SHARED_DRAW) {
                const uint32_t gid0 = a.gc.env_id_base + (uint32_t)base + (uint32_t)lane;
                const U4 draw = sample_block(a.gc.seed, gid0, step);
                U4 draw_next{0, 0, 0, 0};
                if (((gid0 + (uint32_t)(T - 32)) ^ gid0) & ~127u) draw_next = sample_block(a.gc.seed, (gid0 | 127u) + 1u + (gid0 & 31u), step);
                const uint32_t w8[8] = {draw.x, draw.y, draw.z, draw.w, draw_next.x, draw_next.y, draw_next.z, draw_next.w};
                const uint32_t first = gid0 >> 5 & 3u;  // word of the lane's first slot inside its group's block
#pragma unroll
                for (int k = 0; k < NQ; ++k)  // slot k takes word first + k of the two blocks (compile-time indices only: registers)
                    queue[k] = first == 0 ? w8[k & 7] : first == 1 ? w8[(k + 1) & 7] : first == 2 ? w8[(k + 2) & 7] : w8[(k + 3) & 7];
            }
        };
        fill_queue(step_now);
        // The previous tile's store was issued just before this tile's load: waiting for it to be performed and releasing
        // its epoch (a GPU-scope fence) costs about what the load takes, and the warp has nothing else to do until then.
        if (lane == 0 && stored != 0xffffffffu && !a.no_epoch) epoch_publish(&d.epoch[stored], launch + 1u);
        mbar_wait(bar, it & 1u);
#ifdef PKB_TIMING
        wt_lw[it < 2 ? it : 2] += (unsigned long long)(clock64() - wt_last);
#endif
        PKB_WTICK(0);

        // PokerGameEnv mode: the passes below run once per ROUND.  Round 0 plays the caller's action for seat 0 (or, for
        // PokerGameEnv.reset, resets the selected envs); every further round plays one sampled opponent move in the envs
        // whose reference loop has not returned yet (env_advance).  The per-slot control byte lives in act_row.
        // From round 1 on the passes visit only the slots of the PENDING list (built with ballots, like lists E and S).
        uint32_t env_round = 0, n_pend = 0;
        if (ENV && a.env_reset) {
            for (int s = lane; s < T; s += 32) {
                bool pending = false;
                uint32_t vm_now = 0;
                if (base + s < n_envs) {
                    Meta m = unpack_meta(v.meta[s]);
                    if (act_row[s]) {  // Game.reset (game_env.py:23)
                        Seats x{0, 0, 0, 0};
                        ep[s] += 1;
                        reset_env<P>(v, s, x, m, a.gc);
                        store_seats<P>(v, s, x);
                        m.vm = valid_mask<P>(v, s, m.act);
                        v.meta[s] = pack_meta(m);
                        pending = m.act != 0;
                        act_row[s] = (uint8_t)(EC_RESET_LOOP | (pending ? EC_PENDING : 0u));
                    } else
                        act_row[s] = 0;
                    vm_now = valid_mask<P>(v, s, m.act);  // the valid output of envs no round will touch
                }
                out_v[s] = (uint8_t)vm_now;
                out_f[s] = 0;
                const uint32_t vote = __ballot_sync(FULL, pending);
                if (pending) pend[n_pend + (uint32_t)__popc(vote & lanes_below)] = (uint8_t)s;
                n_pend += (uint32_t)__popc(vote);
            }
            env_round = 1;
            __syncwarp();
        } else if (ENV && a.env_ids) {  // continue phase: every slot is an env whose loop had not returned when the last phase stopped
            for (int s = lane; s < T; s += 32) {
                const bool have = base + s < n_envs;
                if (have) {
                    act_row[s] = a.env_ctl_in[base + s];
                    out_f[s] = 0;
                    out_v[s] = (uint8_t)unpack_meta(v.meta[s]).vm;
                    pend[s] = (uint8_t)s;
                }
            }
            n_pend = (uint32_t)(n_envs - base < (size_t)T ? n_envs - base : (size_t)T);
            env_round = 1;
            __syncwarp();
        }
        for (bool env_more = !(ENV && (a.env_reset || a.env_ids)) || n_pend != 0; env_more;) {
        // ---- pass A: every slot
        uint32_t n_e = 0, n_s = 0;  // list lengths (warp-uniform)
        const uint32_t n_live = n_envs - base < (size_t)T ? (uint32_t)(n_envs - base) : (uint32_t)T;  // slots of this tile that are envs
        const uint32_t n_visit = ENV && env_round ? n_pend : (uint32_t)T;
        PKB_UNROLL(PKB_WT_UNROLL_A)
        for (uint32_t v0 = 0; v0 < n_visit; v0 += 32) {
            // every slot of the tile in slot order -- or, in the env mode's later rounds, the slots of the pending list
            const int s = ENV && env_round ? (int)(v0 + lane < n_pend ? pend[v0 + lane] : 0u) : (int)(v0 + lane);
            const size_t env = base + s;
            const bool live = ENV && env_round ? v0 + lane < n_pend : (uint32_t)s < n_live;
            Ctl c{PC_DONE, 0, 0, 0, 0, 0, 0, 0};
            if (live) {
                Seats x = load_seats<P>(v, s);
                Meta m = unpack_meta(v.meta[s]);
                const double hb = high_bet<P>(v, s), credit = v.cr[m.act * T + s];
                // The sampler draws from the mask the previous step (or reset / pkb_refresh_valid) cached in the meta word; a
                // caller's action is checked against get_valid_actions() of the planes as they are now (game.py:648), so
                // that state the caller wrote through the zero-copy view is honoured.
                const uint32_t vm = SAMPLE ? m.vm : valid_mask_from(hb, credit, v.mr[s]);
                int action;
                if (SAMPLE) {
                    const uint32_t word = SHARED_DRAW ? queue[0] : sample_word(a.gc.seed, a.gc.env_id_base + (uint32_t)env, step_now);
                    action = a.use_weights ? pick_weighted(word, vm, a.weights) : pick_uniform(word, vm);
                    act_row[s] = (uint8_t)action;  // pkb_step_random(actions_out): copied out with the tile's outputs
                } else if (ENV && env_round) {  // an opponent's move: only envs whose loop is still running are visited
                    const uint32_t word = sample_word(a.gc.seed, a.gc.env_id_base + (a.env_ids ? a.env_ids[env] : (uint32_t)env),
                                                      a.round_base + a.env_round_offset + env_round - 1u);
                    action = a.use_weights ? pick_weighted(word, vm, a.weights) : pick_uniform(word, vm);
                } else {
                    action = act_row[s];
                    action = action == 255 ? ACTION_SKIP : action;
                }
                step_begin<P, !SAMPLE>(v, s, x, m, c, action, vm, hb, credit, a.gc);
                const bool skipped = !SAMPLE && action == ACTION_SKIP;  // rejected like any out-of-range action, but not an error
                tl.live += skipped ? 0u : 1u;
                tl.invalid += skipped ? 0u : (uint32_t)c.invalid;
                if (c.pc == PC_END_HAND) {  // end_hand's head (game.py:453-482): it is short, and it decides the list
                    ++tl.hands;
                    if (end_hand_begin<P>(v, s, x, c, true)) c.pc = PC_SHOWDOWN;  // step_begin has committed
                }
                if (!c.invalid) store_seats<P>(v, s, x);  // an invalid action changes nothing (game.py:649-651)
                if (c.pc == PC_DONE) {
                    tl.poisoned += (uint32_t)c.poisoned;
                    write_outputs<P>(v, s, c, m, out_f, out_v, skipped);
                }
                v.meta[s] = pack_meta(m);  // after an invalid action m is what it was, with the valid mask recomputed
            }
#pragma unroll
            for (int k = 0; k + 1 < NQ; ++k) queue[k] = queue[k + 1];  // next slot's draw to the head
            const bool in_e = live && c.pc != PC_DONE, in_s = live && c.pc == PC_SHOWDOWN;
            const uint32_t vote_e = __ballot_sync(FULL, in_e), vote_s = __ballot_sync(FULL, in_s);
            if (in_e) list_e[n_e + (uint32_t)__popc(vote_e & lanes_below)] = (uint32_t)s | pack_ctl(c) << 8;
            if (in_s) list_s[n_s + (uint32_t)__popc(vote_s & lanes_below)] = (uint8_t)s;
            n_e += (uint32_t)__popc(vote_e);
            n_s += (uint32_t)__popc(vote_s);
        }
        __syncwarp();
        PKB_WTICK(1);
        // the previous tile's store has had pass A's time to finish: publish it without waiting
        // one look at the next tile's epoch while passes C and D run: the acquire's round trip is off the load's path
        if (lane == 0) handed_over = next < num_tiles && !a.no_epoch && ld_acquire_u32(&d.epoch[next]) == launch;

        // Passes C and D repeat while hands keep ending inside this step (all-in blinds: a slot whose next hand
        // cannot be acted on either); such slots go round again through the SAME code, side by side.
        while (n_e) {
#ifdef PKB_TIMING
            const long long wt_c0 = clock64();
            wt_rounds += 1;
            wt_maxns = n_s > wt_maxns ? n_s : wt_maxns;
#endif
            // ---- pass C: the ranked showdown for the slots in list S, 32 hands at a time, in three phases that each keep
            // the lanes busy with their own kind of work item (pkb_game.cuh): deal per hand, rank per (hand, seat), pots per hand
            for (uint32_t i0 = 0; i0 < n_s; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool have = i < n_s;
                int s = 0;
                Seats x{0, 0, 0, 0};
                if (have) {
                    s = list_s[i];
                    x = load_seats<P>(v, s);
                    ++tl.shows;
                    showdown_deal<P>(v, s, a.gc.seed, a.gc.env_id_base + (ENV && a.env_ids ? a.env_ids[base + s] : (uint32_t)(base + s)), ep[s]);
                }
                if constexpr (P == 2) {  // both seats of every hand: lane pairs
                    __syncwarp();
                    const uint32_t hands = n_s - i0 < 32u ? n_s - i0 : 32u;
                    for (uint32_t j = lane; j < 2u * hands; j += 32) showdown_rank(v, list_s[i0 + (j >> 1)], (int)(j & 1u));
                } else {  // the seats that show their cards (game.py:488), listed hand by hand
                    const uint32_t todo = x.called | x.allin;
                    uint32_t upto = (uint32_t)popc32(todo);  // inclusive prefix sum over the lanes
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y = __shfl_up_sync(FULL, upto, o);
                        upto += lane >= o ? y : 0u;
                    }
                    const uint32_t total = __shfl_sync(FULL, upto, 31);
                    uint32_t at = upto - (uint32_t)popc32(todo);
                    for (uint32_t left = todo; left; left &= left - 1) pairs[at++] = (uint16_t)((uint32_t)s | (uint32_t)(ffs32(left) - 1) << 8);
                    __syncwarp();
                    for (uint32_t j = lane; j < total; j += 32) showdown_rank(v, pairs[j] & 255u, (int)(pairs[j] >> 8));
                }
                __syncwarp();
                if (have) showdown_settle<P>(v, s, x, popc32(x.active | x.called | x.allin));
            }
            __syncwarp();
            PKB_WTICK(2);
#ifdef PKB_TIMING
            if ((unsigned long long)(clock64() - wt_c0) > wt_maxc) wt_maxc = (unsigned long long)(clock64() - wt_c0);
#endif

            // ---- pass D: the rest of end_hand and of the step for the slots in list E.  A slot in which another hand
            // ends is appended to the lists of the next round, in place: at most as many entries are written as were
            // read, and every lane has read its entry before the ballot that precedes the writes.
            uint32_t m_e = 0, m_s = 0;
            for (uint32_t i0 = 0; i0 < n_e; i0 += 32) {
                const uint32_t i = i0 + lane;
                bool again = false, again_s = false;
                uint32_t entry = 0;
                if (i < n_e) {
                    entry = list_e[i];
                    const int s = (int)(entry & 255u);
                    const size_t env = ENV && a.env_ids ? (size_t)a.env_ids[base + s] : base + s;  // the env's index in the handle's planes
                    Seats x = load_seats<P>(v, s);
                    Ctl c = unpack_ctl(entry >> 8);
                    Meta m = unpack_meta(v.meta[s]);
                    if (c.poisoned) {
                        // game.py:473 `assert num_potential_winners > 0` would fire: payoffs were zeroed at :468; the step
                        // reports hand-over (and round-over when end_hand was entered from next_turn) plus PKB_F_POISONED
#pragma unroll
                        for (int q = 0; q < payoff_pitch(P); q += 4) store_f64x4(d.payoffs + env * payoff_pitch(P) + q, 0.0, 0.0, 0.0, 0.0);
                        c.done = c.cont ? (int)((game_over<P>(x) ? F_GAME_OVER : 0u) | F_HAND_OVER | F_ROUND_OVER) : (int)F_HAND_OVER;
                        c.pc = PC_DONE;
                    } else {
                        c.pc = PC_DONE;
                        double pay[P];
                        if (end_hand_finish<P>(v, s, x, m, c, d.payoffs + env * payoff_pitch(P), pay, a.gc)) ep[s] += 1;
#pragma unroll
                        for (int p = 0; p < P; ++p) tl.pay[p] += pay[p];
                        if (c.pc == PC_END_HAND) {  // another hand ends inside this step: its head now, the rest next round
                            ++tl.hands;
                            again = true;
                            again_s = end_hand_begin<P>(v, s, x, c);
                            if (again_s) c.pc = PC_SHOWDOWN;
                            entry = (uint32_t)s | pack_ctl(c) << 8;
                        }
                        store_seats<P>(v, s, x);
                    }
                    if (!again) {
                        const uint32_t f = write_outputs<P>(v, s, c, m, out_f, out_v);
                        tl.games += (f & F_GAME_OVER) ? 1u : 0u;
                        tl.poisoned += (uint32_t)c.poisoned;
                    }
                    v.meta[s] = pack_meta(m);
                }
                const uint32_t vote_e = __ballot_sync(FULL, again), vote_s = __ballot_sync(FULL, again_s);
                if (again) list_e[m_e + (uint32_t)__popc(vote_e & lanes_below)] = entry;
                if (again_s) list_s[m_s + (uint32_t)__popc(vote_s & lanes_below)] = (uint8_t)(entry & 255u);
                m_e += (uint32_t)__popc(vote_e);
                m_s += (uint32_t)__popc(vote_s);
            }
            __syncwarp();
            n_e = m_e;
            n_s = m_s;
        }

            if (!ENV) {
                // pkb_step_random_n: the next Game.step of the same tile, straight from shared memory -- its load and its
                // store are paid once per visit instead of once per step
                if (!SAMPLE || ++rep >= a.reps) break;
                step_now = a.step_idx + rep;
                fill_queue(step_now);
                continue;
            }
            // ---- PokerGameEnv control (envs/game_env.py:20-54): what this round's step means for every env of the tile
            __threadfence_block();  // the payoff plane is written with plain stores by the lanes of pass D and read here
            __syncwarp();
            // the slots that stepped in this round (all of them in round 0, the pending list afterwards); the new pending list
            // is written in place: at most as many entries as were read, each lane having read its own before the ballot
            uint32_t n_next = 0;
            const uint32_t n_stepped = env_round ? n_pend : n_live;
            for (uint32_t v0 = 0; v0 < n_stepped; v0 += 32) {
                const bool have = v0 + lane < n_stepped;
                const int s = env_round ? (int)(have ? pend[v0 + lane] : 0u) : (int)(v0 + lane);
                bool pending = false;
                if (have) {
                    const size_t env = a.env_ids ? (size_t)a.env_ids[base + s] : base + s;  // the env's index in the handle's planes
                    const uint32_t ec_in = env_round == 0 ? (uint32_t)EC_AGENT_STEP : (uint32_t)act_row[s];
                    Seats x = load_seats<P>(v, s);
                    Meta m = unpack_meta(v.meta[s]);
                    const bool cap_hit = a.env_round_offset + env_round + 1u >= (uint32_t)ENV_ROUND_CAP;
                    uint32_t ec = env_advance<P>(ec_in, out_f[s], x, m.act, cap_hit);
                    if (ec & EC_REDO) {  // `if done: self.game.reset()` (game_env.py:27)
                        ep[s] += 1;
                        reset_env<P>(v, s, x, m, a.gc);
                        store_seats<P>(v, s, x);
                        m.vm = valid_mask<P>(v, s, m.act);
                        v.meta[s] = pack_meta(m);
                        out_v[s] = (uint8_t)m.vm;
                        ec = EC_RESET_LOOP | (m.act != 0 && !cap_hit ? EC_PENDING : 0u);
                    }
                    if ((ec & EC_WRITE_REWARD) && a.reward) a.reward[env] = (ec & EC_REWARD) ? d.payoffs[env * payoff_pitch(P)] : 0.0;  // payoffs[agent]: seat 0 leads the env's record
                    if (!(ec & EC_PENDING) && !a.env_reset && a.env_done) a.env_done[env] = (uint8_t)env_out_byte(ec);  // step() returns
                    act_row[s] = (uint8_t)(ec & 255u);
                    pending = (ec & EC_PENDING) != 0;
                }
                const uint32_t vote = __ballot_sync(FULL, pending);
                if (pending) pend[n_next + (uint32_t)__popc(vote & lanes_below)] = (uint8_t)s;
                n_next += (uint32_t)__popc(vote);
            }
            n_pend = n_next;
            env_more = n_pend != 0;
            ++env_round;
            __syncwarp();
            if (env_more && a.env_round_cap && env_round > a.env_round_cap) {
                // this launch's share of the opponent rounds is played: the envs still inside their loops go to the next phase
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(a.strag_count, n_pend);
                at = __shfl_sync(FULL, at, 0);
                for (uint32_t i = lane; i < n_pend; i += 32) {
                    const int s = pend[i];
                    a.strag_ids[at + i] = a.env_ids ? a.env_ids[base + s] : (uint32_t)(base + s);
                    a.strag_ctl[at + i] = act_row[s];
                }
                env_more = false;
            }
        }
        // ---- store: make the generic-proxy writes visible to the async proxy, then push the rows back
        fence_async_smem();
        __syncwarp();
        if (SAMPLE && a.actions_out)
            for (int s = lane; s < T && base + s < n_envs; s += 32) a.actions_out[base + s] = act_row[s];
        if (a.packed) {  // the compact output: FOLD and ALL_IN are always valid (game.py:367), so five mask bits say it all
            for (int s = lane; s < T && base + s < n_envs; s += 32)
                a.packed[base + s] = (uint8_t)((out_f[s] & 7u) | (out_v[s] >> 1 & 31u) << 3);
        }
        if (ENV && a.env_ids) {  // continue phase: the slots are a packed selection of the handle's envs
            for (int s = lane; s < T && base + s < n_envs; s += 32)
                if (a.valid) a.valid[a.env_ids[base + s]] = out_v[s];
        } else if (out_words && base + T <= n_envs) {  // the two output rows: one 32-bit store per lane and 32 slots
#pragma unroll
            for (int w = lane; w < T / 4; w += 32) {
                if (a.flags) reinterpret_cast<uint32_t*>(a.flags + base)[w] = reinterpret_cast<const uint32_t*>(out_f)[w];
                if (a.valid) reinterpret_cast<uint32_t*>(a.valid + base)[w] = reinterpret_cast<const uint32_t*>(out_v)[w];
            }
        } else {
            for (int s = lane; s < T && base + s < n_envs; s += 32) {
                if (a.flags) a.flags[base + s] = out_f[s];
                if (a.valid) a.valid[base + s] = out_v[s];
            }
        }
        PKB_WTICK(3);
        if (lane == 0) {
            tma_store_2d(&a.maps.w64, (int)base, 0, buf_s + (uint32_t)L::OFF_CR);
            tma_store_2d(&a.maps.u32, (int)base, 0, buf_s + (uint32_t)L::OFF_META);
            bulk_commit();
            bulk_wait_read();  // the buffer is about to be refilled
            stored = tile;
        }
        __syncwarp();
        tile = next;
        PKB_WTICK(4);
#ifdef PKB_TIMING
        wt_acc[5] += 1;
#endif
    }
    if (lane == 0 && stored != 0xffffffffu && !a.no_epoch) epoch_publish(&d.epoch[stored], launch + 1u);
#ifdef PKB_TIMING
    {
        const uint32_t gw = blockIdx.x * L::WARPS + warp;
        if (lane == 0 && gw < 4096) {
            unsigned long long gt1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_wrp[gw][0] = wt_gt0;
            g_wrp[gw][1] = gt1;
            g_wrp[gw][2] = smid;
            for (int i = 0; i < 6; ++i) g_wrp[gw][3 + i] = wt_acc[i];
            g_wrp[gw][9] = wt_rounds;
            g_wrp[gw][10] = wt_maxc;
            g_wrp[gw][11] = wt_maxns;
            for (int i = 0; i < 3; ++i) g_wrp[gw][12 + i] = wt_lw[i];
        }
    }
#endif
    if (a.gc.track_stats) reduce_stats<P>(d, tl);
}

// --------------------------------------------------------------------------------------------
// PokerGameEnv.step in phases: the envs whose opponents are still playing when a phase has used up its rounds are copied
// into a scratch arena, packed densely, and copied back after the next phase has stepped them there
// --------------------------------------------------------------------------------------------
template <int P, bool TO_SCRATCH>
__global__ void env_move_kernel(DevState m, DevState sc, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ count)
{
    const uint32_t n = *count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t e = ids[i];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (TO_SCRATCH) {
                sc.cr[p * sc.stride + i] = m.cr[p * m.stride + e];
                sc.bt[p * sc.stride + i] = m.bt[p * m.stride + e];
                sc.pd[p * sc.stride + i] = m.pd[p * m.stride + e];
            } else {
                m.cr[p * m.stride + e] = sc.cr[p * sc.stride + i];
                m.bt[p * m.stride + e] = sc.bt[p * sc.stride + i];
                m.pd[p * m.stride + e] = sc.pd[p * sc.stride + i];
            }
        }
        if (TO_SCRATCH) {
            sc.mr[i] = m.mr[e];
            sc.sm[i] = m.sm[e];
            sc.meta[i] = m.meta[e];
            sc.hand[i] = m.hand[e];
            sc.episode[i] = m.episode[e];
        } else {
            m.mr[e] = sc.mr[i];
            m.sm[e] = sc.sm[i];
            m.meta[e] = sc.meta[i];
            m.hand[e] = sc.hand[i];
            m.episode[e] = sc.episode[i];
        }
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
    Seats x{0, 0, 0, 0};
    reset_env<P>(v, env, x, m, gc);
    store_seats<P>(v, env, x);
    m.vm = valid_mask<P>(v, env, m.act);
    d.meta[env] = pack_meta(m);
}

// recomputes the valid mask cached in the meta word (after the caller has written the state planes)
template <int P>
__global__ void refresh_valid_kernel(DevState d)
{
    const size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n) return;
    const View<0> v = global_view(d);
    Meta m = unpack_meta(d.meta[env]);
    m.vm = valid_mask<P>(v, env, m.act);
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

// Game.player_states (game.py:259) in the reference's form, one PlayerState value per seat: out[p * stride + env]
template <int P>
__global__ void player_states_kernel(DevState d, uint8_t* out)
{
    const size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n) return;
    const Seats x = load_seats<P>(global_view(d), env);
#pragma unroll
    for (int p = 0; p < P; ++p) out[(size_t)p * d.stride + env] = (uint8_t)seat_state(x, p);
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

// The rank-set table of pkb_eval.cuh, built once per device and staged in shared memory by the batch evaluators
__device__ uint32_t g_eval_table[EVAL_TABLE_SIZE];

__global__ void eval_table_init_kernel()
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < EVAL_TABLE_SIZE) g_eval_table[x] = eval_table_entry(x);
}

PKB_D void stage_eval_table(uint32_t* s_tbl)
{
    const uint4* src = reinterpret_cast<const uint4*>(g_eval_table);
    uint4* dst = reinterpret_cast<uint4*>(s_tbl);
    for (int i = threadIdx.x; i < EVAL_TABLE_SIZE / 4; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

constexpr int EVAL_THREADS = 256;

// eval_hand on id masks: 8 B in, 4 B out per hand; two hands per thread per trip (128-bit loads, 64-bit stores)
__global__ void __launch_bounds__(EVAL_THREADS, 6)
eval_masks_kernel(const uint64_t* __restrict__ masks, uint64_t n, uint32_t* __restrict__ out)
{
    __shared__ __align__(16) uint32_t s_tbl[EVAL_TABLE_SIZE];
    stage_eval_table(s_tbl);
    const uint64_t pairs = n / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(masks) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    if (aligned) {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
            const ulonglong2 m = reinterpret_cast<const ulonglong2*>(masks)[i];
            reinterpret_cast<uint2*>(out)[i] = make_uint2(eval_mask52_tbl(m.x, s_tbl), eval_mask52_tbl(m.y, s_tbl));
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) out[n - 1] = eval_mask52_tbl(masks[n - 1], s_tbl);
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = eval_mask52_tbl(masks[i], s_tbl);
    }
}

// ---- the perfect-hash evaluator for seven-card hands (pkb_eval.cuh eval7_hash3): tables built once per device ----
struct EvalHashTables {
    const uint32_t* dense;  // [65536]  slot -> packed value of a rank multiset (256 KB, cache resident)
    const uint32_t* flt;    // [8192]   flush suit's rank set -> top five ranks, straight-flush / wheel-low bits
    const uint32_t* klo;    // [128]    key sum of the low seven ranks of a suit (staged in shared memory)
    const uint32_t* khi;    // [64]     key sum of the high six ranks of a suit (staged in shared memory)
    const uint16_t* disp;   // [30571]  per 256 key sums: the multiplier offset of the displacement hash (61 KB)
};

PKB_D void stage_key_tables(const EvalHashTables& t, uint32_t* s_klo, uint32_t* s_khi)
{
    for (int i = threadIdx.x; i < EVAL_KLO_SIZE; i += blockDim.x) s_klo[i] = t.klo[i];
    for (int i = threadIdx.x; i < EVAL_KHI_SIZE; i += blockDim.x) s_khi[i] = t.khi[i];
    __syncthreads();
}

// seven cards: the hash; anything else (fewer cards, the empty hand): the select evaluator
PKB_D uint32_t eval_any(uint64_t m, const EvalHashTables& t, const uint32_t* s_klo, const uint32_t* s_khi)
{
    return popc64(m) == 7 ? eval7_hash3(m, s_klo, s_khi, t.disp, t.dense, t.flt) : eval_mask52_sel(m);
}

// eval_hand on id masks: 8 B in, 4 B out per hand; two hands per thread per trip (128-bit loads, 64-bit stores)
#ifndef PKB_EVAL_MIN_BLOCKS
#define PKB_EVAL_MIN_BLOCKS 8  // full residency (32 registers): the kernel waits on its three gathers per hand
#endif
#ifndef PKB_EVAL_UNROLL
#define PKB_EVAL_UNROLL 1
#endif
__global__ void __launch_bounds__(EVAL_THREADS, PKB_EVAL_MIN_BLOCKS)
eval_masks_hash_kernel(const uint64_t* __restrict__ masks, uint64_t n, uint32_t* __restrict__ out, EvalHashTables t)
{
    __shared__ uint32_t s_klo[EVAL_KLO_SIZE], s_khi[EVAL_KHI_SIZE];
    stage_key_tables(t, s_klo, s_khi);
    const uint64_t pairs = n / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(masks) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    if (aligned) {
        PKB_UNROLL(PKB_EVAL_UNROLL)
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
            const ulonglong2 m = reinterpret_cast<const ulonglong2*>(masks)[i];
            reinterpret_cast<uint2*>(out)[i] = make_uint2(eval_any(m.x, t, s_klo, s_khi), eval_any(m.y, t, s_klo, s_khi));
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) out[n - 1] = eval_any(masks[n - 1], t, s_klo, s_khi);
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = eval_any(masks[i], t, s_klo, s_khi);
    }
}

// ---- the streaming form (pkb_eval.cuh eval7_fast): every table in shared memory, rare hands queued per warp ----------
//
// One 1024-thread block per SM.  Shared memory holds: the two key tables with one copy per lane (entry x of lane l at
// x * 128 + l * 4: the eight key look-ups of a hand never meet in a bank), the byte-wide first hash level, the 65,536
// value slots as 16-bit class numbers and the class -> packed value table (the seven-card values without a flush are
// ~3,500 distinct words).  A hand costs eight conflict-free and three scattered shared-memory loads and no global
// gather; eval7_hash3's kernel above spent its time in the L1 (three global gathers plus bank conflicts, l1tex 86 %).
// A warp trip ranks 32 x EVAL_FAST_HPT hands (two 128-bit loads and two 64-bit stores per lane).  Hands that the common
// case does not cover (a flush suit, the 5-4-3-2 low run, not seven cards: ~3 % of random hands) leave a wrong word in
// `out`; their index goes to the warp's ring buffer (ballot + popcount prefix), and whenever 32 are waiting the warp
// ranks them together with the select evaluator and overwrites the words -- so the branch-free evaluator always runs on
// full warps instead of inside a divergent branch (in place it costs 2x: 116 G against 235 G evals/s measured).
struct EvalFastTables {
    const uint32_t* klo;      // [128]   fast_klo_entry
    const uint32_t* khi;      // [64]    fast_khi_entry
    const uint8_t* disp;      // [EVAL_FAST_BUCKETS]
    const uint16_t* dense16;  // [EVAL_DENSE_SIZE] slot -> class
    const uint32_t* vals;     // [nvals] class -> packed value
    uint32_t nvals;
};
constexpr int EVAL_FAST_THREADS = 1024, EVAL_FAST_HPT = 4, EVAL_FAST_QSIZE = 256;  // ring >= 31 + 32 * HPT entries
constexpr uint64_t EVAL_FAST_MIN_HANDS = 1ull << 17;  // below this the staging of the tables (~190 KB per block) does not pay
static size_t eval_fast_smem(uint32_t nvals)
{
    return 128 * 32 * 4 + 64 * 32 * 4 + EVAL_FAST_BUCKETS + (EVAL_FAST_THREADS / 32) * EVAL_FAST_QSIZE * 4 + EVAL_DENSE_SIZE * 2 + (size_t)nvals * 4;
}

PKB_D uint32_t lds_off(const void* base, uint32_t off) { return *reinterpret_cast<const uint32_t*>(static_cast<const char*>(base) + off); }

// eval7_fast on the device: byte offsets into the per-lane copies of the key tables
PKB_D uint32_t eval7_fast_dev(uint32_t lo, uint32_t hi, const uint32_t* s_klo, const uint32_t* s_khi, uint32_t lane4, const uint8_t* s_disp,
                              const uint16_t* s_dense16, const uint32_t* s_vals, uint32_t& rare)
{
    const uint32_t mid = __funnelshift_r(lo, hi, 19);  // bits 26.. of the mask at bit 7
    const uint32_t s0 = lds_off(s_klo, ((lo << 7) & 0x3F80u) | lane4) + lds_off(s_khi, (lo & 0x1F80u) | lane4);
    const uint32_t s1 = lds_off(s_klo, ((lo >> 6) & 0x3F80u) | lane4) + lds_off(s_khi, ((lo >> 13) & 0x1F80u) | lane4);
    const uint32_t s2 = lds_off(s_klo, (mid & 0x3F80u) | lane4) + lds_off(s_khi, ((hi << 6) & 0x1F80u) | lane4);
    const uint32_t s3 = lds_off(s_klo, (hi & 0x3F80u) | lane4) + lds_off(s_khi, ((hi >> 7) & 0x1F80u) | lane4);
    const uint32_t sum = s0 + s1 + s2 + s3;
    rare = ((s0 | s1 | s2 | s3) & EVAL_FAST_FLUSH_BIT) | ((sum ^ EVAL_FAST_SEVEN) >> 24);
    return s_vals[s_dense16[fast_slot(sum, s_disp[fast_bucket(sum)])]];
}

// the block's copy of the tables: [128][32] low-part keys | [64][32] high-part keys | first hash level | the warps' rings |
// 16-bit class slots | class -> value
struct EvalFastSmem {
    uint32_t* klo;
    uint32_t* khi;
    uint8_t* disp;
    uint32_t* queue;
    uint16_t* dense16;
    uint32_t* vals;
};
PKB_D EvalFastSmem stage_fast_tables(uint8_t* smem, const EvalFastTables& t)
{
    constexpr int THREADS = EVAL_FAST_THREADS, WARPS = THREADS / 32, QSIZE = EVAL_FAST_QSIZE;
    EvalFastSmem m;
    m.klo = reinterpret_cast<uint32_t*>(smem);
    m.khi = m.klo + 128 * 32;
    m.disp = reinterpret_cast<uint8_t*>(m.khi + 64 * 32);
    m.queue = reinterpret_cast<uint32_t*>(m.disp + EVAL_FAST_BUCKETS);
    m.dense16 = reinterpret_cast<uint16_t*>(m.queue + WARPS * QSIZE);
    m.vals = reinterpret_cast<uint32_t*>(m.dense16 + EVAL_DENSE_SIZE);
    for (int i = threadIdx.x; i < 128 * 32; i += THREADS) m.klo[i] = t.klo[i >> 5];
    for (int i = threadIdx.x; i < 64 * 32; i += THREADS) m.khi[i] = t.khi[i >> 5];
    for (int i = threadIdx.x; i < (int)EVAL_FAST_BUCKETS / 16; i += THREADS) reinterpret_cast<uint4*>(m.disp)[i] = reinterpret_cast<const uint4*>(t.disp)[i];
    for (int i = threadIdx.x; i < (int)EVAL_DENSE_SIZE / 8; i += THREADS) reinterpret_cast<uint4*>(m.dense16)[i] = reinterpret_cast<const uint4*>(t.dense16)[i];
    for (int i = threadIdx.x; i < (int)t.nvals; i += THREADS) m.vals[i] = t.vals[i];
    __syncthreads();
    return m;
}

// ranks the `count` (<= 32) hands whose indices wait at q[head ..] with the select evaluator; one out-of-line copy
__device__ __noinline__ void eval_fast_drain(const uint32_t* q, uint32_t head, uint32_t count, const uint64_t* __restrict__ masks,
                                             uint32_t* __restrict__ out)
{
    __syncwarp();  // the ring's entries, and the words the common case stored for these hands, come first
    const uint32_t lane = threadIdx.x & 31;
    if (lane < count) {
        const uint32_t idx = q[(head + lane) & (EVAL_FAST_QSIZE - 1)];
        out[idx] = eval_mask52_sel(masks[idx]);
    }
}

// masks 16-byte aligned, out 8-byte aligned, n < 2^32
__global__ void __launch_bounds__(EVAL_FAST_THREADS, 1)
eval_masks_fast_kernel(const uint64_t* __restrict__ masks, uint64_t n, uint32_t* __restrict__ out, EvalFastTables t)
{
    constexpr int THREADS = EVAL_FAST_THREADS, WARPS = THREADS / 32, HPT = EVAL_FAST_HPT, QSIZE = EVAL_FAST_QSIZE;
    extern __shared__ __align__(128) uint8_t fast_smem[];
    const EvalFastSmem sm = stage_fast_tables(fast_smem, t);
    const uint32_t *s_klo = sm.klo, *s_khi = sm.khi, *s_vals = sm.vals;
    const uint8_t* s_disp = sm.disp;
    const uint16_t* s_dense16 = sm.dense16;
    uint32_t* s_queue = sm.queue;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, lane4 = lane * 4, lt = (1u << lane) - 1u;
    const uint64_t trips = n / (32 * HPT), wstride = (uint64_t)gridDim.x * WARPS;
    uint32_t head = 0, tail = 0;
    uint32_t* q = s_queue + w * QSIZE;
    const uint64_t first = (uint64_t)blockIdx.x * WARPS + w;
    ulonglong2 nxt[HPT / 2];
    if (first < trips) {
#pragma unroll
        for (int j = 0; j < HPT / 2; ++j) nxt[j] = reinterpret_cast<const ulonglong2*>(masks)[first * (16 * HPT) + lane + 32 * j];
    }
    for (uint64_t trip = first; trip < trips; trip += wstride) {
        const uint64_t pair0 = trip * (16 * HPT) + lane;  // this lane's pairs of hands: pair0 + 32 j
        ulonglong2 m[HPT / 2];
#pragma unroll
        for (int j = 0; j < HPT / 2; ++j) m[j] = nxt[j];
        if (trip + wstride < trips) {  // the next trip's masks are in flight while this one is ranked (+4 %: 335 -> 348 G evals/s)
#pragma unroll
            for (int j = 0; j < HPT / 2; ++j) nxt[j] = reinterpret_cast<const ulonglong2*>(masks)[pair0 + wstride * (16 * HPT) + 32 * j];
        }
        uint32_t r[HPT], rare[HPT];
#pragma unroll
        for (int j = 0; j < HPT / 2; ++j) {
            r[2 * j] = eval7_fast_dev((uint32_t)m[j].x, (uint32_t)(m[j].x >> 32), s_klo, s_khi, lane4, s_disp, s_dense16, s_vals, rare[2 * j]);
            r[2 * j + 1] = eval7_fast_dev((uint32_t)m[j].y, (uint32_t)(m[j].y >> 32), s_klo, s_khi, lane4, s_disp, s_dense16, s_vals, rare[2 * j + 1]);
        }
#pragma unroll
        for (int j = 0; j < HPT / 2; ++j) reinterpret_cast<uint2*>(out)[pair0 + 32 * j] = make_uint2(r[2 * j], r[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < HPT; ++j) {
            const uint32_t bal = __ballot_sync(0xffffffffu, rare[j] != 0);
            if (rare[j]) q[(tail + __popc(bal & lt)) & (QSIZE - 1)] = (uint32_t)(2 * (pair0 + 32 * (j / 2)) + (j & 1));
            tail += __popc(bal);
        }
        while (tail - head >= 32) {
            eval_fast_drain(q, head, 32, masks, out);
            head += 32;
        }
    }
    // the last n % (32 HPT) hands: the block's first warp queues them all
    if (blockIdx.x == 0 && w == 0)
        for (uint64_t i = trips * (32 * HPT) + lane; i < n + lane; i += 32) {  // n + lane: every lane makes the same number of trips
            const uint32_t bal = __ballot_sync(0xffffffffu, i < n);
            if (i < n) q[(tail + __popc(bal & lt)) & (QSIZE - 1)] = (uint32_t)i;
            tail += __popc(bal);
        }
    while (tail != head) {
        const uint32_t count = tail - head < 32 ? tail - head : 32;
        eval_fast_drain(q, head, count, masks, out);
        head += count;
    }
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
                       unsigned long long* sums, EvalHashTables tabs)
{
    __shared__ unsigned long long s_hist[11], s_sum[2];
    __shared__ __align__(16) uint32_t s_tbl[EVAL_TABLE_SIZE];  // k < 7: the rank-set table; k == 7: its first words hold the key tables
    uint32_t* s_klo = s_tbl;
    uint32_t* s_khi = s_tbl + EVAL_KLO_SIZE;
    if (threadIdx.x < 11) s_hist[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_sum[threadIdx.x] = 0;
    if (k == 7) stage_key_tables(tabs, s_klo, s_khi);  // seven cards: the perfect hash; fewer: the rank-set table
    else stage_eval_table(s_tbl);
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = t * EXH_CHUNK;
    unsigned long long cats = 0, sum_a = 0, sum_b = 0;
    if (lo < count) {
        const int todo = (int)(count - lo < EXH_CHUNK ? count - lo : EXH_CHUNK);
        uint64_t m = unrank_colex(first + lo, k);
        for (int j = 0; j < todo; ++j) {
            const uint32_t p = k == 7 ? eval7_hash3(m, s_klo, s_khi, tabs.disp, tabs.dense, tabs.flt) : eval_mask52_tbl(m, s_tbl);
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

// All C(52,7) hands through the streaming evaluator's tables: one 1024-thread block per SM, every lane ranks chunks of
// EXH_CHUNK consecutive hands.  Neighbouring lanes enumerate neighbouring ranges, whose hands share their high cards, so a
// warp's hands are rare (a flush suit, ...) nearly all together or not at all: the select evaluator runs in place here,
// no queue (with random hands the same branch halves the throughput, eval_masks_fast_kernel).
__global__ void __launch_bounds__(EVAL_FAST_THREADS, 1)
eval_exhaustive7_kernel(uint64_t first, uint64_t count, uint32_t* __restrict__ out, unsigned long long* hist, unsigned long long* sums,
                        EvalFastTables t)
{
    constexpr int WARPS = EVAL_FAST_THREADS / 32;
    extern __shared__ __align__(128) uint8_t fast_smem[];
    __shared__ unsigned long long s_hist[11], s_sum[2];
    if (threadIdx.x < 11) s_hist[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_sum[threadIdx.x] = 0;
    const EvalFastSmem sm = stage_fast_tables(fast_smem, t);
    const uint32_t lane = threadIdx.x & 31, lane4 = lane * 4;
    const uint64_t chunks = (count + EXH_CHUNK - 1) / EXH_CHUNK, wstride = (uint64_t)gridDim.x * WARPS * 32;
    unsigned long long sum_a = 0, sum_b = 0;
    for (uint64_t cb = ((uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * 32; cb < chunks; cb += wstride) {
        const uint64_t lo = (cb + lane) * EXH_CHUNK;
        const int todo = lo < count ? (int)(count - lo < EXH_CHUNK ? count - lo : EXH_CHUNK) : 0;
        unsigned long long cats = 0;  // 6-bit counters per category: at most EXH_CHUNK = 32 per chunk
        if (todo) {
            uint64_t m = unrank_colex(first + lo, 7);
            for (int j = 0; j < todo; ++j) {
                uint32_t rare;
                uint32_t p = eval7_fast_dev((uint32_t)m, (uint32_t)(m >> 32), sm.klo, sm.khi, lane4, sm.disp, sm.dense16, sm.vals, rare);
                if (rare) p = eval_mask52_sel(m);
                if (out) out[lo + j] = p;
                cats += 1ull << (6 * (p >> 20));
                sum_a += p;
                sum_b += (unsigned long long)p * (m % 1000003ull);
                const uint64_t c = m & (0 - m), r = m + c;  // next mask with the same popcount (Gosper)
                m = (((r ^ m) >> 2) >> (__ffsll((long long)c) - 1)) | r;
            }
        }
        for (int cat = 1; cat <= 9; ++cat) {  // seven cards: categories 1..9
            const uint32_t w = __reduce_add_sync(0xffffffffu, (uint32_t)(cats >> (6 * cat)) & 63u);
            if (lane == 0 && w) atomicAdd(&s_hist[cat], (unsigned long long)w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_a += __shfl_xor_sync(0xffffffffu, sum_a, o);
        sum_b += __shfl_xor_sync(0xffffffffu, sum_b, o);
    }
    if (lane == 0) {
        atomicAdd(&s_sum[0], sum_a);
        atomicAdd(&s_sum[1], sum_b);
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
    TileMaps maps;
    pkb_config cfg;
    DevState d;
    GameConst gc;
    void* arena;
    size_t arena_bytes;
    uint32_t* d_weights;
    int sm_count;
    // The stream of the handle's last launch.  Step launches of one handle hand tiles over through epoch words and derive
    // their launch index from the order in which their blocks start, so they must start in submission order: a call on a
    // DIFFERENT stream is first ordered after everything submitted to the previous one.
    cudaStream_t last_stream;
    bool has_last;
    cudaEvent_t order_event;
    bool pipelined;      // step launches may overlap their predecessor's tail (programmatic dependent launch)
    bool last_was_step;  // the handle's previous launch was a step launch (on last_stream)
    unsigned long long last_capture;  // id of the stream capture that launch was recorded into (0: launched live)
    // PokerGameEnv.step in phases (env_step_phases): scratch arena for the envs whose loop outlasts a phase
    int64_t env_phase_min;   // handles with at least this many envs step their env mode in phases (< 0: never)
    void* scratch;
    DevState sd;             // the scratch arena's planes (payoffs and statistics are the handle's own)
    TileMaps smaps;
    uint32_t* strag_count[2];
    uint32_t* strag_ids[2];
    uint8_t* strag_ctl[2];
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

// orders the handle's next launch (on stream s) after its previous ones when the caller has switched streams
static cudaError_t order_on_stream(pkb_env* h, cudaStream_t s, bool step_launch = false)
{
    cudaError_t e = cudaSuccess;
    if (h->has_last && h->last_stream != s) {
        e = cudaEventRecord(h->order_event, h->last_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, h->order_event, 0);
        if (e != cudaSuccess) {  // the old stream is gone: everything on the device finishes first
            (void)cudaGetLastError();
            e = cudaDeviceSynchronize();
        }
    }
    h->last_was_step = step_launch && h->last_was_step && h->has_last && h->last_stream == s;
    h->last_stream = s;
    h->has_last = true;
    return e;
}

// Does the step launch about to be issued on `s` follow, at run time, the handle's previous step launch?  Only then may
// it skip griddepcontrol.wait.  A launch recorded into a CUDA graph follows its recorded predecessor only if that one
// went into the SAME capture; the first launch of a capture, and the first live launch after one, wait.
static uint32_t step_must_wait(pkb_env* h, cudaStream_t s)
{
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo(s, &status, &id) != cudaSuccess) {
        (void)cudaGetLastError();
        status = cudaStreamCaptureStatusNone;
    }
    const unsigned long long now = status == cudaStreamCaptureStatusActive ? id : 0ull;
    const bool same_context = now == h->last_capture;
    h->last_capture = now;
    return h->last_was_step && same_context ? 0u : 1u;
}

static inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// One launch per step: a persistent grid of as many blocks as fit the device at once.
template <class K>
static cudaError_t resident_blocks(K kernel, int threads, size_t smem, int device, int* cache)
{
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    if (cache[device]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // the tiles want (almost) all of the SM's 228 KB as shared memory; without this the driver may pick a smaller
    // carve-out, fewer blocks become resident than the persistent grid assumes, and the rest run as a second wave
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    cache[device] = nb;
    return cudaSuccess;
}

// Every step launch of a handle -- caller's actions or device sampler -- uses the same grid (the launch index is derived
// from it on the device), so the resident-block count is the smaller of the two kernel variants'.
template <int P>
static cudaError_t step_grid(int device, int sm_count, uint64_t n, unsigned* grid, uint32_t* num_tiles)
{
    using L = WarpTile<P>;
    static int per_sm_t[64] = {0}, per_sm_f[64] = {0};
    cudaError_t e = resident_blocks(step_warp_kernel<P, true>, L::THREADS, L::BYTES, device, per_sm_t);
    if (e == cudaSuccess) e = resident_blocks(step_warp_kernel<P, false>, L::THREADS, L::BYTES, device, per_sm_f);
    if (e != cudaSuccess) return e;
    int per_sm = per_sm_t[device] < per_sm_f[device] ? per_sm_t[device] : per_sm_f[device];
    per_sm = per_sm < L::MIN_BLOCKS ? per_sm : L::MIN_BLOCKS;
    static const bool verbose = getenv("PKB_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "pkb: step kernel P=%d warp tile=%d smem=%zu blocks/SM=%d (compiled for %d)\n", P, L::T, (size_t)L::BYTES, per_sm,
                L::MIN_BLOCKS);
    *num_tiles = (uint32_t)((n + L::T - 1) / L::T);
    const unsigned resident = (unsigned)(per_sm * sm_count);
    const unsigned want = (*num_tiles + L::WARPS - 1) / L::WARPS;
    *grid = want < resident ? want : resident;
    return cudaSuccess;
}

// One launch per step: a persistent grid of as many blocks as fit the device at once, launched with programmatic stream
// serialization so that it may start while the previous step launch of the stream is still draining (tiles are handed
// over one by one through their epoch words, see epoch_wait).
template <int P, bool SAMPLE, bool ENV = false>
static cudaError_t launch_step(StepArgs& a, cudaStream_t s, int device, int sm_count, bool pipelined)
{
    using L = WarpTile<P>;
    unsigned grid = 0;
    cudaError_t e = step_grid<P>(device, sm_count, a.d.n, &grid, &a.num_tiles);
    if (e != cudaSuccess) return e;
    if (ENV) {  // same grid as every other step launch of the handle (the launch index is derived from it); if fewer of
                // this variant's blocks fit an SM, the rest simply start later
        static int per_sm_env[64] = {0};
        e = resident_blocks(step_warp_kernel<P, SAMPLE, ENV>, L::THREADS, L::BYTES, device, per_sm_env);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(L::THREADS);
    cfg.dynamicSmemBytes = L::BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pipelined ? 1 : 0;  // pkb_set_pipelining(h, 0): plain stream order between step launches
    return cudaLaunchKernelEx(&cfg, step_warp_kernel<P, SAMPLE, ENV>, a);
}

// ---- TMA descriptors (driver entry point fetched through the runtime: no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// [rows][stride] plane group of `elem`-byte elements, box = [rows][tile]
static bool make_plane_map(CUtensorMap* m, CUtensorMapDataType type, size_t elem, void* base, uint64_t stride, uint32_t rows, uint32_t tile)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {stride, rows};
    const cuuint64_t strides[1] = {stride * elem};
    const cuuint32_t box[2] = {tile, rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, type, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int P>
static bool make_tile_maps(TileMaps& maps, const DevState& d)
{
    using L = WarpTile<P>;
    return make_plane_map(&maps.w64, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, d.cr, d.stride, L::W64_ROWS, L::T) &&
           make_plane_map(&maps.u32, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, d.meta, d.stride, L::U32_ROWS, L::T);
}

extern "C" {

#ifdef PKB_TIMING
// debug builds only: the per-warp timeline of the last step launch (profiles/warp_timing.py)
int pkb_debug_warps(unsigned long long* out) { return cudaMemcpyFromSymbol(out, g_wrp, sizeof(unsigned long long) * 4096 * 16) == cudaSuccess ? 0 : -1; }
#endif

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
    const uint64_t n = cfg->num_envs, stride = (n + STRIDE_ALIGN - 1) / STRIDE_ALIGN * STRIDE_ALIGN;
    // arena: 8-byte planes (f64 money, seat masks) | u32 planes | stats
    const size_t n_w64 = (size_t)(3 * P + 2 + payoff_pitch(P)) * stride, n_u32 = (size_t)3 * stride;  // payoffs: one padded record per env
    const size_t off_u32 = n_w64 * 8, off_stats = (off_u32 + n_u32 * 4 + 255) / 256 * 256;
    const size_t off_seq = (off_stats + STAT_COUNT * 8 + MAX_PLAYERS * 8 + NUM_MOVES * 4 + 4 + 255) / 256 * 256;
    const size_t n_epoch = stride / 32;  // one word per tile; tiles hold at least 32 envs
    h->arena_bytes = off_seq + 256 + n_epoch * 4;
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
    d.mr = f + (size_t)3 * P * stride;  // right after pending_bets: credits .. seat masks are one 2-D TMA tensor of 3P+2 rows
    d.sm = reinterpret_cast<uint64_t*>(f + (size_t)(3 * P + 1) * stride);
    d.payoffs = f + (size_t)(3 * P + 2) * stride;
    uint32_t* u = reinterpret_cast<uint32_t*>(base + off_u32);
    d.meta = u;
    d.hand = u + stride;
    d.episode = u + 2 * stride;
    d.stats = reinterpret_cast<unsigned long long*>(base + off_stats);
    d.payoff_sum = reinterpret_cast<double*>(base + off_stats + STAT_COUNT * 8);
    h->d_weights = reinterpret_cast<uint32_t*>(base + off_stats + STAT_COUNT * 8 + MAX_PLAYERS * 8);
    d.seq = reinterpret_cast<unsigned long long*>(base + off_seq);
    d.epoch = reinterpret_cast<uint32_t*>(base + off_seq + 256);
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
    e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
    if (e != cudaSuccess) {
        cudaFree(h->arena);
        delete h;
        return fail("pkb_create: cudaDeviceGetAttribute", e);
    }
    bool maps_ok = false;
    PKB_DISPATCH_P(P, fill_fresh_state<P>(gc); maps_ok = make_tile_maps<P>(h->maps, d));
    if (!maps_ok) {
        cudaFree(h->arena);
        delete h;
        return fail("pkb_create: cuTensorMapEncodeTiled failed");
    }
    h->env_phase_min = 16384;
    h->scratch = nullptr;
    h->has_last = false;
    h->last_stream = nullptr;
    h->pipelined = getenv("PKB_NO_PDL") == nullptr;
    h->last_was_step = false;
    e = cudaEventCreateWithFlags(&h->order_event, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaFree(h->arena);
        delete h;
        return fail("pkb_create: cudaEventCreate", e);
    }
    *out = h;
    return 0;
}

int pkb_destroy(pkb_env* h)
{
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->arena);  // waits for the handle's launches
    cudaFree(h->scratch);
    cudaEventDestroy(h->order_event);
    delete h;
    return 0;
}

int pkb_set_reset_dealer(pkb_env* h, int32_t dealer)
{
    if (!h) return fail("pkb_set_reset_dealer: null handle");
    if (dealer < 0 || dealer >= h->cfg.num_players) return fail("pkb_set_reset_dealer: dealer out of range");
    h->cfg.dealer = dealer;
    h->gc.dealer = dealer;
    PKB_DISPATCH_P(h->cfg.num_players, fill_fresh_state<P>(h->gc));
    return 0;
}

int pkb_set_pipelining(pkb_env* h, int32_t enabled)
{
    if (!h) return fail("pkb_set_pipelining: null handle");
    h->pipelined = enabled != 0;
    return 0;
}

int pkb_refresh_valid(pkb_env* h, void* stream)
{
    if (!h) return fail("pkb_refresh_valid: null handle");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_CUDA(order_on_stream(h, s));
    PKB_DISPATCH_P(h->cfg.num_players, (refresh_valid_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_reset(pkb_env* h, const uint8_t* env_mask, void* stream)
{
    if (!h) return fail("pkb_reset: null handle");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_CUDA(order_on_stream(h, s));
    PKB_DISPATCH_P(h->cfg.num_players, (reset_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, h->gc, env_mask)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

static int step_common(pkb_env* h, bool sample, uint32_t step_idx, const uint32_t* weights, const int32_t* actions,
                       const uint8_t* actions_u8, int32_t* actions_out, uint8_t* flags, uint8_t* valid, void* stream,
                       const uint8_t* actions_u4 = nullptr, uint8_t* packed = nullptr, uint32_t reps = 1)
{
    if (!h) return fail("pkb_step: null handle");
    if (!sample && !actions && !actions_u8 && !actions_u4) return fail("pkb_step: null actions");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    StepArgs a{};  // every field the call does not set is zero (null pointers, no phases, tile hand-over on)
    a.maps = h->maps;
    a.d = h->d;
    a.gc = h->gc;
    a.actions = actions;
    a.actions_u8 = actions_u8;
    a.actions_u4 = actions_u4;
    a.packed = packed;
    a.actions_out = actions_out;
    a.env_reset = 0;
    a.round_base = 0;
    a.env_mask = nullptr;
    a.reward = nullptr;
    a.env_done = nullptr;
    a.flags = flags;
    a.valid = valid;
    a.step_idx = step_idx;
    a.reps = reps;
    a.use_weights = weights ? 1u : 0u;
    for (int i = 0; i < NUM_MOVES; ++i) a.weights[i] = weights ? weights[i] : 1u;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_CUDA(order_on_stream(h, s, true));
    // Only the fused random-agent step is launched ahead of its predecessor's completion: all it reads is the handle's
    // own state.  A step with the caller's actions reads a buffer some other kernel or copy of the stream has just
    // produced, so it keeps plain stream order.  (Other work the caller puts on the stream BETWEEN two pkb_step_random
    // calls must not write the handle's state; pkb_refresh_valid / pkb_reset between them are fine.)
    const bool ahead = sample && h->pipelined;
    a.wait_prev = step_must_wait(h, s);
    cudaError_t e = cudaSuccess;
    PKB_DISPATCH_P(h->cfg.num_players,
                   e = sample ? launch_step<P, true>(a, s, h->cfg.device, h->sm_count, ahead)
                              : launch_step<P, false>(a, s, h->cfg.device, h->sm_count, ahead));
    if (e != cudaSuccess) return fail("pkb_step: launch", e);
    h->last_was_step = true;
    return 0;
}

int pkb_step(pkb_env* h, const int32_t* actions, uint8_t* flags, uint8_t* valid, void* stream)
{
    return step_common(h, false, 0, nullptr, actions, nullptr, nullptr, flags, valid, stream);
}

int pkb_step_u8(pkb_env* h, const uint8_t* actions, uint8_t* flags, uint8_t* valid, void* stream)
{
    return step_common(h, false, 0, nullptr, nullptr, actions, nullptr, flags, valid, stream);
}

// opponent rounds per phase of PokerGameEnv.step (0: until every loop has returned; ENV_ROUND_CAP still applies)
constexpr int ENV_MAX_PHASES = 8;
struct EnvPhases {
    int count;
    uint32_t rounds[ENV_MAX_PHASES];
};
static const EnvPhases& env_phases()
{
    static const EnvPhases ph = [] {
        EnvPhases p{5, {3, 6, 12, 24, 0}};  // measured on 2^23 four-player tables: 6,12,32,0 10.45 ms per learner step; 3,6,12,24,0 9.87; 4,0 10.9
        if (const char* e = getenv("PKB_ENV_PHASE_ROUNDS")) {  // tuning: comma-separated rounds per phase, the last one 0
            p.count = 0;
            for (const char* q = e; *q && p.count < ENV_MAX_PHASES;) {
                p.rounds[p.count++] = (uint32_t)strtoul(q, nullptr, 10);
                q = strchr(q, ',');
                if (!q) break;
                ++q;
            }
            if (p.count == 0 || p.rounds[p.count - 1] != 0) p = EnvPhases{5, {3, 6, 12, 24, 0}};
        }
        return p;
    }();
    return ph;
}

// the scratch arena of a handle: the planes a step launch moves (same layout, same stride), two straggler lists
static int env_scratch_create(pkb_env* h)
{
    const int P = h->cfg.num_players;
    const uint64_t stride = h->d.stride;
    const size_t n_w64 = (size_t)(3 * P + 2) * stride, n_u32 = (size_t)3 * stride;
    const size_t off_u32 = n_w64 * 8, off_seq = (off_u32 + n_u32 * 4 + 255) / 256 * 256;
    const size_t off_cnt = off_seq + 256, off_ids = off_cnt + 256, off_ctl = off_ids + 2 * stride * 4;
    const size_t bytes = off_ctl + 2 * stride;
    void* mem = nullptr;
    cudaError_t e = cudaMalloc(&mem, bytes);
    if (e != cudaSuccess) return fail("pkb_env_step: cudaMalloc (scratch arena)", e);
    e = cudaMemset(mem, 0, bytes);
    if (e != cudaSuccess) {
        cudaFree(mem);
        return fail("pkb_env_step: cudaMemset (scratch arena)", e);
    }
    char* base = static_cast<char*>(mem);
    double* f = reinterpret_cast<double*>(base);
    DevState sd = h->d;  // payoffs, statistics: the handle's own
    sd.cr = f;
    sd.bt = f + (size_t)P * stride;
    sd.pd = f + (size_t)2 * P * stride;
    sd.mr = f + (size_t)3 * P * stride;
    sd.sm = reinterpret_cast<uint64_t*>(f + (size_t)(3 * P + 1) * stride);
    uint32_t* u = reinterpret_cast<uint32_t*>(base + off_u32);
    sd.meta = u;
    sd.hand = u + stride;
    sd.episode = u + 2 * stride;
    sd.seq = reinterpret_cast<unsigned long long*>(base + off_seq);
    sd.epoch = nullptr;  // scratch launches keep plain stream order (StepArgs::no_epoch)
    bool maps_ok = false;
    PKB_DISPATCH_P(P, maps_ok = make_tile_maps<P>(h->smaps, sd));
    if (!maps_ok) {
        cudaFree(mem);
        return fail("pkb_env_step: cuTensorMapEncodeTiled failed (scratch arena)");
    }
    h->sd = sd;
    for (int i = 0; i < 2; ++i) {
        h->strag_count[i] = reinterpret_cast<uint32_t*>(base + off_cnt) + 16 * i;
        h->strag_ids[i] = reinterpret_cast<uint32_t*>(base + off_ids) + (size_t)i * stride;
        h->strag_ctl[i] = reinterpret_cast<uint8_t*>(base + off_ctl) + (size_t)i * stride;
    }
    h->scratch = mem;
    return 0;
}

// PokerGameEnv.step / reset for every env in one launch (the kernel's env mode)
static int env_common(pkb_env* h, bool reset, const int32_t* actions, const uint8_t* env_mask, uint32_t round_base,
                      const uint32_t* weights, double* reward, uint8_t* done, uint8_t* valid, void* stream)
{
    if (!h) return fail("pkb_env: null handle");
    if (!reset && !actions) return fail("pkb_env_step: null actions");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    StepArgs a{};  // every field the call does not set is zero (null pointers, no phases, tile hand-over on)
    a.maps = h->maps;
    a.d = h->d;
    a.gc = h->gc;
    a.gc.auto_reset = 0;  // the env decides about resets itself (game_env.py:27 inside reset(), the caller after step())
    a.actions = actions;
    a.actions_u8 = nullptr;
    a.actions_u4 = nullptr;
    a.packed = nullptr;
    a.actions_out = nullptr;
    a.flags = nullptr;
    a.valid = valid;
    a.step_idx = 0;
    a.use_weights = weights ? 1u : 0u;
    for (int i = 0; i < NUM_MOVES; ++i) a.weights[i] = weights ? weights[i] : 1u;
    a.env_reset = reset ? 1u : 0u;
    a.round_base = round_base;
    a.env_mask = env_mask;
    a.reward = reward;
    a.env_done = done;
    a.env_round_cap = a.env_round_offset = a.no_epoch = 0;
    a.env_ids = nullptr;
    a.n_dev = nullptr;
    a.env_ctl_in = nullptr;
    a.strag_count = nullptr;
    a.strag_ids = nullptr;
    a.strag_ctl = nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    bool phases = !reset && h->env_phase_min >= 0 && h->d.n >= (uint64_t)h->env_phase_min;
    if (phases && !h->scratch) {
        // the scratch arena is allocated by the first phased call; a call that is being recorded into a CUDA graph cannot
        // allocate, so it runs as one launch (same results) -- call pkb_env_step once before capturing to get the phases
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) phases = false;
        else if (env_scratch_create(h) != 0) return -1;
    }
    PKB_CUDA(order_on_stream(h, s, true));
    a.wait_prev = step_must_wait(h, s);
    cudaError_t e = cudaSuccess;
    if (!phases) {
        PKB_DISPATCH_P(h->cfg.num_players, e = (launch_step<P, false, true>(a, s, h->cfg.device, h->sm_count, false)));
        if (e != cudaSuccess) return fail("pkb_env: launch", e);
        h->last_was_step = true;
        return 0;
    }
    // PokerGameEnv.step in phases.  A table whose agent has folded plays its hand out among the opponents, and one whose
    // agent is broke plays until the game is over, inside the same call (envs/game_env.py:41-52): 5 % of the tables need
    // more than 6 opponent moves, 1 % more than 28, and a 128-table tile would go round ~35 times for its last few tables
    // (measured: 8.9 of 32 lanes).  So the launch over the handle's planes stops after the first phase's opponent
    // rounds; the envs still inside their loops are copied into a scratch arena, packed densely, stepped there by the
    // next phase, and copied back.  An env's j-th opponent move draws with the same sampler index in whichever phase
    // it falls, so the result does not depend on the phases.
    PKB_CUDA(cudaMemsetAsync(h->strag_count[0], 0, 4, s));
    const EnvPhases& ph = env_phases();
    a.env_round_cap = ph.rounds[0];
    a.strag_count = h->strag_count[0];
    a.strag_ids = h->strag_ids[0];
    a.strag_ctl = h->strag_ctl[0];
    PKB_DISPATCH_P(h->cfg.num_players, e = (launch_step<P, false, true>(a, s, h->cfg.device, h->sm_count, false)));
    if (e != cudaSuccess) return fail("pkb_env: launch", e);
    uint32_t offset = ph.rounds[0];
    for (int k = 1; k < ph.count; ++k) {
        const int cur = (k - 1) & 1, nxt = k & 1;
        StepArgs b = a;
        b.maps = h->smaps;
        b.d = h->sd;
        b.actions = nullptr;
        b.env_round_cap = ph.rounds[k];
        b.env_round_offset = offset;
        b.no_epoch = 1;
        b.wait_prev = 1;
        b.env_ids = h->strag_ids[cur];
        b.n_dev = h->strag_count[cur];
        b.env_ctl_in = h->strag_ctl[cur];
        b.strag_count = h->strag_count[nxt];
        b.strag_ids = h->strag_ids[nxt];
        b.strag_ctl = h->strag_ctl[nxt];
        PKB_CUDA(cudaMemsetAsync(h->strag_count[nxt], 0, 4, s));
        const unsigned mg = (unsigned)(h->sm_count * 4);
        PKB_DISPATCH_P(h->cfg.num_players, (env_move_kernel<P, true><<<mg, 256, 0, s>>>(h->d, h->sd, b.env_ids, b.n_dev)));
        PKB_DISPATCH_P(h->cfg.num_players, e = (launch_step<P, false, true>(b, s, h->cfg.device, h->sm_count, false)));
        if (e != cudaSuccess) return fail("pkb_env: phase launch", e);
        PKB_DISPATCH_P(h->cfg.num_players, (env_move_kernel<P, false><<<mg, 256, 0, s>>>(h->d, h->sd, b.env_ids, b.n_dev)));
        PKB_CUDA(cudaGetLastError());
        offset += ph.rounds[k];
    }
    h->last_was_step = false;  // the last launch wrote the planes outside the tile hand-over: the next step launch waits for it
    return 0;
}

int pkb_set_env_phases(pkb_env* h, int64_t min_envs)
{
    if (!h) return fail("pkb_set_env_phases: null handle");
    h->env_phase_min = min_envs;
    return 0;
}

int pkb_env_step(pkb_env* h, const int32_t* actions, uint32_t round_base, const uint32_t* weights, double* reward, uint8_t* done,
                 uint8_t* valid, void* stream)
{
    return env_common(h, false, actions, nullptr, round_base, weights, reward, done, valid, stream);
}

int pkb_env_reset(pkb_env* h, const uint8_t* env_mask, uint32_t round_base, const uint32_t* weights, uint8_t* valid, void* stream)
{
    return env_common(h, true, nullptr, env_mask, round_base, weights, nullptr, nullptr, valid, stream);
}

int pkb_step_packed(pkb_env* h, const uint8_t* actions_u4, uint8_t* packed_out, void* stream)
{
    if (!actions_u4 || !packed_out) return fail("pkb_step_packed: null argument");
    return step_common(h, false, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream, actions_u4, packed_out);
}

int pkb_step_random(pkb_env* h, uint32_t step_idx, const uint32_t* weights, int32_t* actions_out, uint8_t* flags,
                    uint8_t* valid, void* stream)
{
    return step_common(h, true, step_idx, weights, nullptr, nullptr, actions_out, flags, valid, stream);
}

int pkb_step_random_n(pkb_env* h, uint32_t step_idx, uint32_t n_steps, const uint32_t* weights, int32_t* actions_out, uint8_t* flags,
                      uint8_t* valid, void* stream)
{
    if (n_steps == 0) return 0;
    return step_common(h, true, step_idx, weights, nullptr, nullptr, actions_out, flags, valid, stream, nullptr, nullptr, n_steps);
}

int pkb_sample_actions(pkb_env* h, uint32_t step_idx, const uint32_t* weights, int32_t* actions_out, void* stream)
{
    if (!h || !actions_out) return fail("pkb_sample_actions: null argument");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_CUDA(order_on_stream(h, s));
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
    PKB_CUDA(order_on_stream(h, s));
    PKB_DISPATCH_P(h->cfg.num_players, (valid_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, player, valid)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_game_over(pkb_env* h, uint8_t* out, void* stream)
{
    if (!h || !out) return fail("pkb_game_over: null argument");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_CUDA(order_on_stream(h, s));
    PKB_DISPATCH_P(h->cfg.num_players, (game_over_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, out)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_player_states(pkb_env* h, uint8_t* out, void* stream)
{
    if (!h || !out) return fail("pkb_player_states: null argument");
    PKB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PKB_CUDA(order_on_stream(h, s));
    PKB_DISPATCH_P(h->cfg.num_players, (player_states_kernel<P><<<blocks_for(h->d.n, 256), 256, 0, s>>>(h->d, out)));
    PKB_CUDA(cudaGetLastError());
    return 0;
}

int pkb_state(pkb_env* h, pkb_state_view* out)
{
    if (!h || !out) return fail("pkb_state: null argument");
    out->num_envs = h->d.n;
    out->stride = h->d.stride;
    out->num_players = h->cfg.num_players;
    out->payoff_pitch = payoff_pitch(h->cfg.num_players);
    out->credits = h->d.cr;
    out->bets = h->d.bt;
    out->pending_bets = h->d.pd;
    out->payoffs = h->d.payoffs;
    out->minimum_raise_value = h->d.mr;
    out->seat_masks = h->d.sm;
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
    PKB_CUDA(order_on_stream(h, s));
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
    PKB_CUDA(order_on_stream(h, s));
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
    PKB_CUDA(order_on_stream(h, s));
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
    PKB_CUDA(order_on_stream(h, static_cast<cudaStream_t>(stream)));
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

// builds the evaluator's rank-set table on the current device the first time a batch evaluator runs there
static std::mutex g_tables_mutex;  // the per-device table caches below are shared by every host thread

static cudaError_t ensure_eval_table(cudaStream_t s)
{
    static bool ready[64] = {false};
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!ready[dev]) {
        eval_table_init_kernel<<<EVAL_TABLE_SIZE / 256, 256, 0, s>>>();
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // built once: a first use from another stream must see all of it
        if (e == cudaSuccess) ready[dev] = true;
    }
    return e;
}

// builds the perfect-hash evaluator's tables on the current device the first time a batch evaluator runs there
// (host: 49,205 rank multisets through the product's own evaluator, ~30 ms; then one 31 MB upload)
static cudaError_t ensure_eval_hash_tables(EvalHashTables* out)
{
    static EvalHashTables tabs[64];
    static bool ready[64] = {false};
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!ready[dev]) {
        // host: key sum -> value for the 49,205 rank multisets, packed into the dense table by the displacement hash
        uint32_t* nf = static_cast<uint32_t*>(malloc((size_t)EVAL_NF_SIZE * 4));
        const size_t words = (size_t)EVAL_DENSE_SIZE + 8192 + EVAL_KLO_SIZE + EVAL_KHI_SIZE + (EVAL_DISP_SIZE + 1) / 2;  // dense | flt | klo | khi | disp (u16)
        uint32_t* host = static_cast<uint32_t*>(malloc(words * 4));
        if (!nf || !host) {
            free(nf);
            free(host);
            return cudaErrorMemoryAllocation;
        }
        uint32_t* h_flt = host + EVAL_DENSE_SIZE;
        uint32_t* h_klo = h_flt + 8192;
        uint32_t* h_khi = h_klo + EVAL_KLO_SIZE;
        uint16_t* h_disp = reinterpret_cast<uint16_t*>(h_khi + EVAL_KHI_SIZE);
        const bool built = eval_nf_table_fill(nf) == 49205u && eval_dense_table_fill(nf, h_disp, host);
        free(nf);
        if (!built) {
            free(host);
            return cudaErrorUnknown;  // the rank keys collide or a bucket found no multiplier: cannot happen with the constants in pkb_eval.cuh
        }
        for (uint32_t x = 0; x < 8192u; ++x) h_flt[x] = flush_table_entry(rot_ranks(x));  // all three indexed by RAW suit slices
        for (uint32_t x = 0; x < (uint32_t)EVAL_KLO_SIZE; ++x) h_klo[x] = raw_key_sum(x);
        for (uint32_t x = 0; x < (uint32_t)EVAL_KHI_SIZE; ++x) h_khi[x] = raw_key_sum(x << 7);
        uint32_t* dptr = nullptr;
        e = cudaMalloc(&dptr, words * 4);
        if (e == cudaSuccess) e = cudaMemcpy(dptr, host, words * 4, cudaMemcpyHostToDevice);
        free(host);
        if (e != cudaSuccess) return e;
        uint32_t* d_flt = dptr + EVAL_DENSE_SIZE;
        tabs[dev] = EvalHashTables{dptr, d_flt, d_flt + 8192, d_flt + 8192 + EVAL_KLO_SIZE,
                                   reinterpret_cast<const uint16_t*>(d_flt + 8192 + EVAL_KLO_SIZE + EVAL_KHI_SIZE)};
        ready[dev] = true;
    }
    *out = tabs[dev];
    return cudaSuccess;
}

// builds the streaming evaluator's tables on the current device the first time it runs there (host: ~50 ms)
static cudaError_t ensure_eval_fast_tables(EvalFastTables* out)
{
    static EvalFastTables tabs[64];
    static bool ready[64] = {false};
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!ready[dev]) {
        uint32_t* nf = static_cast<uint32_t*>(malloc((size_t)EVAL_NF_SIZE * 4));
        uint32_t* dense = static_cast<uint32_t*>(malloc((size_t)EVAL_DENSE_SIZE * 4 * 2));  // values, then a sorted copy
        // one upload: klo | khi | vals (<= EVAL_DENSE_SIZE) | dense16 | disp
        const size_t words = 128 + 64 + EVAL_DENSE_SIZE + EVAL_DENSE_SIZE / 2 + EVAL_FAST_BUCKETS / 4;
        uint32_t* host = static_cast<uint32_t*>(malloc(words * 4));
        if (!nf || !dense || !host) {
            free(nf);
            free(dense);
            free(host);
            return cudaErrorMemoryAllocation;
        }
        uint32_t* h_klo = host;
        uint32_t* h_khi = h_klo + 128;
        uint32_t* h_vals = h_khi + 64;
        uint16_t* h_dense16 = reinterpret_cast<uint16_t*>(h_vals + EVAL_DENSE_SIZE);
        uint8_t* h_disp = reinterpret_cast<uint8_t*>(h_dense16 + EVAL_DENSE_SIZE);
        const bool built = eval_nf_table_fill(nf, true) == 49205u && eval_fast_table_fill(nf, h_disp, dense);
        free(nf);
        uint32_t nvals = 0;
        if (built) {  // classes: the distinct words of the value table, ascending (the empty slots' 0 is class 0)
            uint32_t* sorted = dense + EVAL_DENSE_SIZE;
            memcpy(sorted, dense, (size_t)EVAL_DENSE_SIZE * 4);
            std::sort(sorted, sorted + EVAL_DENSE_SIZE);
            nvals = (uint32_t)(std::unique(sorted, sorted + EVAL_DENSE_SIZE) - sorted);
            memcpy(h_vals, sorted, (size_t)nvals * 4);
            for (uint32_t i = 0; i < EVAL_DENSE_SIZE; ++i) h_dense16[i] = (uint16_t)(std::lower_bound(sorted, sorted + nvals, dense[i]) - sorted);
        }
        free(dense);
        if (!built || nvals > 8192u) {  // 8192 classes = 32 KB is what the kernel's shared-memory budget allows (there are ~3,500)
            free(host);
            return cudaErrorUnknown;
        }
        for (uint32_t x = 0; x < 128u; ++x) h_klo[x] = fast_klo_entry(x);
        for (uint32_t x = 0; x < 64u; ++x) h_khi[x] = fast_khi_entry(x);
        uint32_t* dptr = nullptr;
        e = cudaMalloc(&dptr, words * 4);
        if (e == cudaSuccess) e = cudaMemcpy(dptr, host, words * 4, cudaMemcpyHostToDevice);
        free(host);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(eval_masks_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eval_fast_smem(nvals));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(eval_exhaustive7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eval_fast_smem(nvals));
        if (e != cudaSuccess) {
            cudaFree(dptr);
            return e;
        }
        uint32_t* d_vals = dptr + 128 + 64;
        const uint16_t* d_dense16 = reinterpret_cast<const uint16_t*>(d_vals + EVAL_DENSE_SIZE);
        tabs[dev] = EvalFastTables{dptr, dptr + 128, reinterpret_cast<const uint8_t*>(d_dense16 + EVAL_DENSE_SIZE), d_dense16, d_vals, nvals};
        ready[dev] = true;
    }
    *out = tabs[dev];
    return cudaSuccess;
}

int pkb_eval_masks(const uint64_t* masks, uint64_t n, uint32_t* out, void* stream)
{
    if (n == 0) return 0;
    if (!out || !masks) return fail("pkb_eval_masks: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static const bool use_table = getenv("PKB_EVAL_IMPL") && !strcmp(getenv("PKB_EVAL_IMPL"), "table");
    static const bool use_hash3 = getenv("PKB_EVAL_IMPL") && !strcmp(getenv("PKB_EVAL_IMPL"), "hash3");  // round 2's first kernel, for comparison
    const uint64_t want = (n / 2 + EVAL_THREADS - 1) / EVAL_THREADS;
    if (use_table) {  // the rank-set-table evaluator (comparison): six 256-thread blocks per SM, the 32 KB table each
        PKB_CUDA(ensure_eval_table(s));
        const unsigned grid = (unsigned)(want < 148ull * 6 ? (want ? want : 1) : 148ull * 6);
        eval_masks_kernel<<<grid, EVAL_THREADS, 0, s>>>(masks, n, out);
    } else if (n >= EVAL_FAST_MIN_HANDS && !use_hash3 && (reinterpret_cast<uintptr_t>(masks) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
        // the streaming evaluator: one block per SM, every table in shared memory; ring entries are 32-bit hand indices
        EvalFastTables t;
        PKB_CUDA(ensure_eval_fast_tables(&t));
        int dev = 0, sms = 0;
        PKB_CUDA(cudaGetDevice(&dev));
        PKB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const uint64_t chunk = 1ull << 31;
        for (uint64_t lo = 0; lo < n; lo += chunk) {
            const uint64_t cnt = n - lo < chunk ? n - lo : chunk;
            eval_masks_fast_kernel<<<(unsigned)sms, EVAL_FAST_THREADS, eval_fast_smem(t.nvals), s>>>(masks + lo, cnt, out + lo, t);
        }
    } else {  // small batches and unaligned arrays
        EvalHashTables t;
        PKB_CUDA(ensure_eval_hash_tables(&t));
        const unsigned grid = (unsigned)(want < 148ull * PKB_EVAL_MIN_BLOCKS ? (want ? want : 1) : 148ull * PKB_EVAL_MIN_BLOCKS);
        eval_masks_hash_kernel<<<grid, EVAL_THREADS, 0, s>>>(masks, n, out, t);
    }
    PKB_CUDA(cudaGetLastError());
    return 0;
}

static cudaError_t ensure_binom()
{
    static bool done = false;  // per process; constant memory is per device context
    static int done_device = -1;
    std::lock_guard<std::mutex> lock(g_tables_mutex);
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
    PKB_CUDA(ensure_eval_table(s));
    EvalHashTables t;
    PKB_CUDA(ensure_eval_hash_tables(&t));
    unsigned long long* d_acc = nullptr;
    PKB_CUDA(cudaMalloc(&d_acc, 13 * 8));
    cudaError_t e = cudaMemsetAsync(d_acc, 0, 13 * 8, s);
    static const bool use_hash3 = getenv("PKB_EVAL_IMPL") && !strcmp(getenv("PKB_EVAL_IMPL"), "hash3");
    if (e == cudaSuccess && count) {
        const uint64_t threads = (count + EXH_CHUNK - 1) / EXH_CHUNK;
        if (k == 7 && !use_hash3 && threads >= 4096) {  // the streaming evaluator's tables, one block per SM
            EvalFastTables ft;
            e = ensure_eval_fast_tables(&ft);
            int dev = 0, sms = 0;
            if (e == cudaSuccess) e = cudaGetDevice(&dev);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (e == cudaSuccess) {
                eval_exhaustive7_kernel<<<(unsigned)sms, EVAL_FAST_THREADS, eval_fast_smem(ft.nvals), s>>>(first, count, out, d_acc, d_acc + 11, ft);
                e = cudaGetLastError();
            }
        } else {
            eval_exhaustive_kernel<<<blocks_for(threads, 256), 256, 0, s>>>(k, first, count, out, d_acc, d_acc + 11, t);
            e = cudaGetLastError();
        }
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
