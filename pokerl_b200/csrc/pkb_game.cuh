// pkb_game.cuh -- the reference Game state machine (pokerl/game.py) as device functions over a
// structure-of-arrays view of N lockstep environments.
//
// Everything here follows the reference line by line, including its non-standard behaviours
// (SURVEY.md Appendix C): money is IEEE double, products are never fused into adds
// (game.py:370-371, 675-678), sums follow numpy's add.reduce order, and the winner scan keeps
// the judger.py:148 behaviour.
//
// Representation: the per-seat PlayerState values are kept -- in HBM, in the tile and in registers -- as four
// disjoint seat bit masks (struct Seats; FOLDED = in none of them), so every np.sum /
// np.argmax / np.roll over player_states of the reference (game.py:320, 337, 421, 471, 567-572,
// 598-599, 683) is a popcount / find-first-set instead of a loop over seats.  Money planes are
// addressed through a View: element (seat p, slot s) of a plane is plane[p * ld + s]; the same
// code runs on a shared-memory staging tile and directly on the HBM planes.
#pragma once
#include "pkb_common.cuh"
#include "pkb_eval.cuh"
#include "pkb_philox.cuh"

namespace pkb {

constexpr uint32_t F_GAME_OVER = 1u, F_HAND_OVER = 2u, F_ROUND_OVER = 4u, F_SKIPPED = 32u, F_INVALID = 64u, F_POISONED = 128u;
constexpr int ACTION_SKIP = -2147483647 - 1;
constexpr int NEXT_TURN_CAP = 64;
constexpr int MAX_PLAYERS = 10;

// per-handle constants, passed by value to every kernel (Game(**config), game.py:246-251)
// Game.payoffs (game.py:264) is stored env-major: one record of payoff_pitch(P) doubles per env, P payoffs and zero padding
// up to a multiple of four.  Payoffs are written only by the steps that end a hand (a third of the heads-up steps), from
// scattered envs: with seat-major planes every such env touched one 32-byte sector per seat and left most of each sector
// unwritten -- the L2 had to fetch the rest from HBM before it could write the sector back (27 MB of the 186 MB a heads-up
// launch moved, and 4.9 % of its time, measured with the stores compiled out).  A record is whole sectors, written by
// 256-bit stores: nothing is fetched, and one store per env replaces one per seat.
PKB_HD constexpr int payoff_pitch(int P) { return (P + 3) & ~3; }
PKB_D void store_f64x4(double* p, double a, double b, double c, double d)  // p 32-byte aligned
{
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}

struct GameConst {
    double start_credits[MAX_PLAYERS];
    double big_blind, small_blind;
    uint64_t seed;
    uint32_t env_id_base;
    int32_t dealer;
    int32_t auto_reset;
    int32_t track_stats;
    // The state right after Game.reset() (game.py:397-412) -- a function of the configuration only, worked out once on the
    // host (fill_fresh_state) so that the auto-reset of a finished game is a copy, not a second setup_hand.
    // credits = start_credits, bets = 0, hand = 1.
    double fresh_pending[MAX_PLAYERS];
    double fresh_min_raise;
    uint32_t fresh_meta, fresh_active, fresh_called, fresh_allin;
};

template <int LD>
struct View {
    double* cr;      // credits          [P][ld]
    double* bt;      // bets             [P][ld]
    double* pd;      // pending_bets     [P][ld]
    double* mr;      // minimum_raise    [ld]
    uint64_t* sm;    // player_states    [ld]     four 16-bit seat masks per env (read/written only by load_seats / store_seats)
    uint32_t* meta;  // packed indices   [ld]
    uint32_t* hand;  // hand counter     [ld]
    size_t ld_rt;
    PKB_HDM size_t ld() const { return LD ? (size_t)LD : ld_rt; }
};

// turn | dealer<<3 | small_blind_idx<<8 | big_blind_idx<<13 | active_player<<18 | valid<<23
// `valid` caches get_valid_actions() of the active player (7 bits): every step ends by working it out for its outputs, and
// the next step validates the action against it instead of recomputing it from the money planes.
struct Meta {
    int turn, dealer, sb, bb, act;
    uint32_t vm;
};
PKB_HD Meta unpack_meta(uint32_t w)
{
    return Meta{(int)(w & 7u), (int)(w >> 3 & 31u), (int)(w >> 8 & 31u), (int)(w >> 13 & 31u), (int)(w >> 18 & 31u), w >> 23 & 127u};
}
PKB_HD uint32_t pack_meta(const Meta& m)
{
    return (uint32_t)m.turn | (uint32_t)m.dealer << 3 | (uint32_t)m.sb << 8 | (uint32_t)m.bb << 13 | (uint32_t)m.act << 18 | m.vm << 23;
}

// ---- seat-state masks (enums.py:130-134: FOLDED 0, ACTIVE 1, CALLED 2, ALL_IN 3, BROKEN 4) -------
struct Seats {
    uint32_t active, called, allin, broken;  // bit p = seat p; disjoint; FOLDED = none of them
};

template <int P>
PKB_HD constexpr uint32_t all_seats() { return (1u << P) - 1u; }

// Storage form of player_states: ACTIVE | CALLED << 16 | ALL_IN << 32 | BROKEN << 48, bit p of a field = seat p
// (pkb_player_states / GameBatch.player_states give the reference's one value per seat).
template <int P, class V>
PKB_HD Seats load_seats(const V& v, size_t s)
{
    const uint64_t w = v.sm[s];
    return Seats{(uint32_t)w & 0xffffu, (uint32_t)(w >> 16) & 0xffffu, (uint32_t)(w >> 32) & 0xffffu, (uint32_t)(w >> 48)};
}

template <int P, class V>
PKB_HD void store_seats(V& v, size_t s, const Seats& x)
{
    v.sm[s] = (uint64_t)(x.active | x.called << 16) | (uint64_t)(x.allin | x.broken << 16) << 32;
}

// one seat's PlayerState value (enums.py:130-134) out of the masks
PKB_HD uint32_t seat_state(const Seats& x, int p)
{
    return (x.active >> p & 1u) * ST_ACTIVE + (x.called >> p & 1u) * ST_CALLED + (x.allin >> p & 1u) * ST_ALL_IN +
           (x.broken >> p & 1u) * ST_BROKEN;
}

// player_states[p] = state
PKB_HD void set_state(Seats& x, int p, int state)
{
    const uint32_t bit = 1u << p, keep = ~bit;
    x.active = (x.active & keep) | (state == ST_ACTIVE ? bit : 0u);
    x.called = (x.called & keep) | (state == ST_CALLED ? bit : 0u);
    x.allin = (x.allin & keep) | (state == ST_ALL_IN ? bit : 0u);
    x.broken = (x.broken & keep) | (state == ST_BROKEN ? bit : 0u);
}

// first set bit of `mask` at or after seat `idx` in cyclic order; mask != 0, idx < 32
PKB_HD int first_from(uint32_t mask, int idx)
{
    const uint32_t hi = mask >> idx << idx;
    return ffs32(hi ? hi : mask) - 1;
}
// the same for a table of P seats; heads-up tables need no bit search: seat idx if it is in the mask, else the other one
template <int P>
PKB_HD int first_from_p(uint32_t mask, int idx)
{
    if constexpr (P == 2) return (mask >> idx & 1u) ? idx : idx ^ 1;
    else return first_from(mask, idx);
}

// numpy float64 add.reduce order (pairwise_sum): n < 8 left to right from 0.0; n >= 8 eight lanes,
// combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail
template <int P>
PKB_D double np_sum(const double (&a)[P])
{
    if constexpr (P < 8) {
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i) r = r + a[i];
        return r;
    } else {
        double r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
#pragma unroll
        for (int i = 8; i < P; ++i) r = r + a[i];
        return r;
    }
}

template <int P, class V>
PKB_D double sum_plane(const V& v, const double* plane, size_t s)
{
    double a[P];
#pragma unroll
    for (int p = 0; p < P; ++p) a[p] = plane[p * v.ld() + s];
    return np_sum<P>(a);
}

// Game.get_first_playing, game.py:334-337; idx in [0, 2P)
template <int P>
PKB_HD int first_playing(const Seats& x, int idx)
{
    if (idx >= P) idx -= P;
    const uint32_t alive = all_seats<P>() & ~x.broken;
    if constexpr (P == 2) return ((alive >> idx & 1u) || !alive) ? idx : idx ^ 1;
    return alive ? first_from(alive, idx) : idx;  // argmax of all-False is 0
}

// Game.high_bet, game.py:286-290
template <int P, class V>
PKB_D double high_bet(const V& v, size_t s)
{
    double m = v.pd[s];
#pragma unroll
    for (int p = 1; p < P; ++p) {
        const double x = v.pd[p * v.ld() + s];
        if (x > m) m = x;
    }
    return m;
}

// Game.game_over, game.py:316-320
template <int P>
PKB_D bool game_over(const Seats& x)
{
    return popc32(all_seats<P>() & ~x.broken) == 1;
}

// Game.get_valid_actions, game.py:339-383 -> bit a set iff move a is valid
PKB_D uint32_t valid_mask_from(double hb, double credit, double mr)
{
    const double room = credit - hb;
    uint32_t m = 1u << MV_FOLD | 1u << MV_ALL_IN;  // :367
    const double f[3] = {0.1, 0.25, 0.5};
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // :370-372, product and sum rounded separately
        const double rv = __dmul_rn(f[i], room);
        m |= (uint32_t)(rv > mr && __dadd_rn(hb, rv) < credit) << (MV_RAISE + i);
    }
    m |= (uint32_t)(hb == 0.0) << MV_CHECK;   // :375
    m |= (uint32_t)(hb < credit) << MV_CALL;  // :376
    return m;
}

template <int P, class V>
PKB_D uint32_t valid_mask(const V& v, size_t s, int player)
{
    return valid_mask_from(high_bet<P>(v, s), v.cr[player * v.ld() + s], v.mr[s]);
}

// Game.setup_hand, game.py:414-451.  The deck is not materialised: hand `hand` of this episode is
// a pure function of (seed, env, episode, hand) and is regenerated where it is needed.
template <int P, class V>
PKB_HD void setup_hand(V& v, size_t s, Seats& x, Meta& m, const GameConst& gc)
{
    const size_t ld = v.ld();
    v.hand[s] += 1;
    m.turn = 0;
    x.active = all_seats<P>() & ~x.broken;  // :421
    x.called = x.allin = 0;
    if constexpr (P == 2) {
        // :432-435 for two seats.  first_playing(idx) is seat idx (mod 2) when both seats or neither are alive, else the
        // one alive seat whatever idx: so dealer, small blind, big blind and first actor alternate d^1, d, d^1, d, or all
        // four are the alive seat.
        const uint32_t alive = 3u & ~x.broken;
        const bool by_index = alive == 3u || alive == 0u;
        const int only = alive == 1u ? 0 : 1, d = m.dealer & 1;
        m.dealer = by_index ? d ^ 1 : only;
        m.sb = by_index ? d : only;
        m.bb = by_index ? d ^ 1 : only;
        m.act = by_index ? d : only;
    } else {
        m.dealer = first_playing<P>(x, m.dealer + 1);  // :432
        m.sb = first_playing<P>(x, m.dealer + 1);      // :433
        m.bb = first_playing<P>(x, m.sb + 1);          // :434
        m.act = first_playing<P>(x, m.bb + 1);         // :435
    }
    set_state(x, m.bb, ST_CALLED);                 // :441
    double hb = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {  // :438-440 (the small blind is written last, so it wins when sb == bb), :444-446
        v.bt[p * ld + s] = 0.0;
        double b = p == m.sb ? gc.small_blind : p == m.bb ? gc.big_blind : 0.0;
        const double c = v.cr[p * ld + s];
        if (b > c) set_state(x, p, ST_ALL_IN);
        if (!(b <= c)) b = c;
        v.pd[p * ld + s] = b;
        hb = (p == 0 || b > hb) ? b : hb;
    }
    v.mr[s] = hb;
}

// Game.reset, game.py:397-412 (the episode counter is bumped by the caller)
template <int P, class V>
PKB_HD void reset_env(V& v, size_t s, Seats& x, Meta& m, const GameConst& gc)
{
    const size_t ld = v.ld();
    m.dealer = gc.dealer;  // :403
    m.act = 0;
    v.hand[s] = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) v.cr[p * ld + s] = gc.start_credits[p];  // :408
    x = Seats{all_seats<P>(), 0, 0, 0};                                  // :409
    setup_hand<P>(v, s, x, m, gc);
}

// host side: fills the fresh_* fields of gc by running Game.reset on one scratch table
template <int P>
inline void fill_fresh_state(GameConst& gc)
{
    double cr[P], bt[P], pd[P], mr[1] = {0.0};
    uint64_t sm[1] = {0};
    uint32_t meta[1] = {0}, hand[1] = {0};
    View<0> v{cr, bt, pd, mr, sm, meta, hand, 1};
    Seats x{0, 0, 0, 0};
    Meta m = unpack_meta(0);
    reset_env<P>(v, 0, x, m, gc);
    for (int p = 0; p < MAX_PLAYERS; ++p) gc.fresh_pending[p] = p < P ? pd[p] : 0.0;
    gc.fresh_min_raise = mr[0];
    gc.fresh_meta = pack_meta(m);
    gc.fresh_active = x.active;
    gc.fresh_called = x.called;
    gc.fresh_allin = x.allin;
}

// game.py:457-461 and 554-558
template <int P, class V>
PKB_D void commit(V& v, size_t s)
{
    const size_t ld = v.ld();
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const double b = v.pd[p * ld + s];
        v.bt[p * ld + s] += b;
        v.cr[p * ld + s] -= b;
        v.pd[p * ld + s] = 0.0;
    }
    v.mr[s] = 0.0;
}

// The betting part of Game.step, game.py:656-696 (the action is already known to be valid), written
// as selects so that the 32 envs of a warp do not diverge on the move.  `hb` = high_bet, `credit` =
// credits[p] (both already read for the valid mask).
template <int P, class V>
PKB_D void apply_action(V& v, size_t s, Seats& x, int p, int action, double hb, double credit, const GameConst& gc)
{
    const bool bets = action >= MV_CALL;                                       // :662 (FOLD :656, CHECK :659 bet nothing)
    const double base = hb >= gc.big_blind ? hb : gc.big_blind;                // :665 max(): first argument on ties
    const double f = action == MV_RAISE ? 0.1 : action == MV_RAISE + 1 ? 0.25 : 0.5;
    const double raised = __dadd_rn(base, __dmul_rn(credit - base, f));        // :673-678, three rounded operations
    const double bet = action == MV_ALL_IN ? credit : action >= MV_RAISE ? raised : base;  // :669-671
    if (bets && bet > hb) {  // :680-687 every CALLED seat becomes ACTIVE again, the actor keeps its new state
        x.active |= x.called;
        x.called = 0;
        v.mr[s] = bet - hb;
    }
    set_state(x, p, action == MV_FOLD ? ST_FOLDED : action == MV_ALL_IN ? ST_ALL_IN : ST_CALLED);  // :657, :660, :667, :671
    if (bets) v.pd[p * v.ld() + s] = bet;  // :696
}

// ---- the resumable part of Game.next_player / next_turn / end_hand ------------------------------
//
// A step is run as a small automaton so that the expensive, rarely taken branches (end_hand and
// the showdown inside it) can be executed by densely packed threads instead of diverging warps.
// Control word, one per parked env:
//   bits 0-1 pc (0 done, 1 needs end_hand, 2 needs showdown), bit 2 cont (1: end_hand was entered
//   from next_turn and the seat scan resumes afterwards, game.py:609-611; 0: from the fold path,
//   game.py:616-619), bits 3-7 current_player of the scan, bits 8-10 the `done` tuple so far,
//   bits 11-17 next_turn() calls made, bit 18 poisoned, bit 19 invalid action,
//   bits 20-23 number of potential winners (hand-over from end_hand_begin to the showdown)
struct Ctl {
    int pc, cont, cur, done, nt, poisoned, invalid, npw;
};
enum : int { PC_DONE = 0, PC_END_HAND = 1, PC_SHOWDOWN = 2 };
PKB_D Ctl unpack_ctl(uint32_t w)
{
    return Ctl{(int)(w & 3u), (int)(w >> 2 & 1u), (int)(w >> 3 & 31u), (int)(w >> 8 & 7u), (int)(w >> 11 & 127u),
               (int)(w >> 18 & 1u), (int)(w >> 19 & 1u), (int)(w >> 20 & 15u)};
}
PKB_D uint32_t pack_ctl(const Ctl& c)
{
    return (uint32_t)c.pc | (uint32_t)c.cont << 2 | (uint32_t)c.cur << 3 | (uint32_t)c.done << 8 | (uint32_t)c.nt << 11 |
           (uint32_t)c.poisoned << 18 | (uint32_t)c.invalid << 19 | (uint32_t)c.npw << 20;
}

// seats act, act+1, .., cur in cyclic order (both < P)
template <int P>
PKB_D uint32_t cyclic_interval(int act, int cur)
{
    if constexpr (P == 2) return act == cur ? 1u << act : 3u;
    const uint32_t ge_act = all_seats<P>() & ~((1u << act) - 1u), le_cur = (2u << cur) - 1u;
    return act <= cur ? (ge_act & le_cur) : (ge_act | le_cur);
}

// One Game.next_turn (game.py:541-576) up to the point where it would call end_hand (turn == 4);
// returns false when the step must stop here (loop cap, or end_hand is due).
template <int P, class V>
PKB_D bool next_turn_once(V& v, size_t s, Seats& x, Meta& m, Ctl& c, bool committed = false)
{
    m.act = c.cur;
    if (c.nt >= NEXT_TURN_CAP) {  // the reference has no termination guarantee here
        c.poisoned = 1;
        return false;
    }
    ++c.nt;
    if (!committed) commit<P>(v, s);  // :554-558 (step_begin has done it already for the step's first next_turn)
    m.turn += 1;      // :561
    if (m.turn == 4) {  // :563-565 -> end_hand, then `return game_over, True, True`
        c.pc = PC_END_HAND;
        c.cont = 1;
        return false;
    }
    if (popc32(x.called) > 1) {  // :567-572
        x.active |= x.called;
        x.called = 0;
    }
    m.act = first_playing<P>(x, m.dealer + 1);  // :575
    c.done = (int)F_ROUND_OVER;                 // :576
    return true;
}

// The `while` of Game.next_player (game.py:607-611) with next_turn inlined.  The reference walks seat
// by seat from active_player until it finds an ACTIVE seat, and calls next_turn when it arrives back
// at the seat that just acted (`current_player`) without finding one: that is "first ACTIVE seat of
// the cyclic interval [active_player .. current_player], else next_turn".
// After ONE next_turn nothing that the search looks at can change any more: pending is zero (further
// commits add 0.0), CALLED seats were merged into ACTIVE (or there is at most one, which never
// merges), and the search restarts from the same seat.  So if the search fails again, every further
// next_turn fails the same way until turn == 4 ("streets skipped"): the
// remaining streets are run out in closed form.
// the part of seat_scan after its search has found no ACTIVE seat (c.pc == PC_DONE on entry); `committed`: the caller
// has already done the commit of the first next_turn
template <int P, class V>
PKB_D void seat_scan_miss(V& v, size_t s, Seats& x, Meta& m, Ctl& c, bool committed)
{
    if (!next_turn_once<P>(v, s, x, m, c, committed)) return;
    uint32_t cand = x.active & cyclic_interval<P>(m.act, c.cur);
    if (!cand) {
        const int more = 4 - m.turn;  // next_turn calls until turn == 4
        if (c.nt + more <= NEXT_TURN_CAP) {
            c.nt += more;
            m.turn = 4;
            m.act = c.cur;
            c.pc = PC_END_HAND;
            c.cont = 1;
            return;
        }
        do {  // the loop cap falls inside the run-out: step through it call by call
            if (!next_turn_once<P>(v, s, x, m, c)) return;
            cand = x.active & cyclic_interval<P>(m.act, c.cur);
        } while (!cand);
    }
    m.act = first_from_p<P>(cand, m.act);
}

template <int P, class V>
PKB_D void seat_scan(V& v, size_t s, Seats& x, Meta& m, Ctl& c)
{
    c.pc = PC_DONE;
    const uint32_t cand = x.active & cyclic_interval<P>(m.act, c.cur);
    if (!cand) {
        seat_scan_miss<P>(v, s, x, m, c, false);
        return;
    }
    m.act = first_from_p<P>(cand, m.act);
}

// Game.step up to the first end_hand (game.py:648-699 and 597-605).  `vmask`, `hb`, `credit`: the
// actor's valid mask and the high bet / stack it was computed from.
// CHECKED = false: the action was drawn from `vmask` on the device (pick_uniform / pick_weighted return a valid move by
// construction), so the validation is left out.
template <int P, bool CHECKED = true, class V>
PKB_D void step_begin(V& v, size_t s, Seats& x, Meta& m, Ctl& c, int action, uint32_t vmask, double hb, double credit,
                      const GameConst& gc)
{
    c = Ctl{PC_DONE, 0, 0, 0, 0, 0, 0, 0};
    if (CHECKED && (action < 0 || action >= NUM_MOVES || !(vmask >> action & 1u))) {  // :648-651, before any mutation
        c.invalid = 1;
        return;
    }
    apply_action<P>(v, s, x, m.act, action, hb, credit, gc);
    const bool contested = popc32(x.active | x.called | x.allin) > 1;  // :598
    uint32_t cand = 0;
    if (contested) {  // :599-605
        c.cur = m.act;
        m.act = m.act + 1 == P ? 0 : m.act + 1;
        cand = x.active & cyclic_interval<P>(m.act, c.cur);
    }
    // Both ways on from here start with the same commit -- next_turn (nobody left to act, :554-558) and end_hand after
    // a fold (:457-461) -- so it is done once, for whichever lanes need it, instead of at two places with few lanes each.
    if (!cand) commit<P>(v, s);
    if (!contested) {  // :616-619; end_hand_begin(.., committed = true)
        c.pc = PC_END_HAND;
        c.cont = 0;
    } else if (cand)
        m.act = first_from_p<P>(cand, m.act);
    else
        seat_scan_miss<P>(v, s, x, m, c, true);
}

// Game.end_hand, first part (game.py:453-482): commit, count the potential winners, settle a
// last-stand pot.  Returns true when the hand needs the ranked showdown (game.py:483-528).
// From here until setup_hand the PENDING plane carries self.payoffs (it is all zero after commit and
// is rewritten by the next setup_hand), so no per-seat register array crosses a phase boundary.
template <int P, class V>
PKB_D bool end_hand_begin(V& v, size_t s, const Seats& x, Ctl& c, bool committed = false)
{
    const size_t ld = v.ld();
    // :457-461; pending (= payoffs from here on) is 0.0 afterwards, :468.  When end_hand is entered from
    // next_turn (c.cont) that call has just committed: pending and the minimum raise are already zero.
    if (!c.cont && !committed) commit<P>(v, s);
    const uint32_t contenders = x.active | x.called | x.allin;  // :471 not BROKEN, not FOLDED
    const int n = popc32(contenders);
    c.npw = n;
    if (n == 0) {  // :473 -- AssertionError in the reference
        c.poisoned = 1;
        return false;
    }
    if (n == 1) {  // :475-482
        const int first = ffs32(contenders) - 1;
        const double pot = sum_plane<P>(v, v.bt, s);
        v.pd[first * ld + s] = pot;
        v.cr[first * ld + s] += pot;
        return false;
    }
    return true;
}

// Game.end_hand, ranked showdown (game.py:483-528), in three phases so that each one runs with every lane busy on its own
// kind of work item:
//   showdown_deal    one lane per hand: replays the deal and parks the board / hole-card masks in the hand's own tile rows
//   showdown_rank    one lane per (hand, seat): eval_hand(board + hole) of one seat (judger.py:7-99)
//   showdown_settle  one lane per hand: the side-pot loop (game.py:496-528)
// Scratch: after end_hand's commit the PENDING row of every seat and the minimum raise are 0.0 and stay unread until
// showdown_settle writes the payoffs into PENDING (minimum raise is rewritten by setup_hand), so the 8-byte words
// pd[p] hold seat p's hole-card mask, then its packed rank, and mr holds the board mask.
template <int P, class V>
PKB_D void showdown_deal(V& v, size_t s, uint64_t seed, uint32_t env_id, uint32_t episode)
{
    const size_t ld = v.ld();
    uint64_t* pd = reinterpret_cast<uint64_t*>(v.pd);
    uint64_t board = 0;
    if constexpr (5 + 2 * P <= 13) {  // small tables: interleaved Philox blocks, sorted draw chain
        uint32_t id[5 + 2 * P];
        deal_ids<5 + 2 * P>(seed, env_id, episode, v.hand[s], id);
        board = 1ull << id[0] | 1ull << id[1] | 1ull << id[2] | 1ull << id[3] | 1ull << id[4];
#pragma unroll
        for (int p = 0; p < P; ++p) pd[p * ld + s] = 1ull << id[5 + 2 * p] | 1ull << id[6 + 2 * p];
    } else {
        // larger tables: one Philox block and its four draws per trip of a rolled loop (straight-line code inside, so the
        // warp's lanes stay together: the sequential Dealer's out-of-line calls ran lane group by lane group)
        constexpr int N = 5 + 2 * P, NB = (N + 3) / 4;
        const uint32_t hand = v.hand[s];
        uint64_t free_mask = (1ull << 52) - 1, pair = 0;
#pragma unroll 1
        for (int b = 0; b < NB; ++b) {
            const U4 blk = philox4x32_10_inl(U4{env_id, hand, episode, (uint32_t)b}, (uint32_t)seed, (uint32_t)(seed >> 32));
            const uint32_t w[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * b + k;
                if (i < N) {
                    const uint64_t bit = 1ull << select_bit_inl(free_mask, mulhi32(w[k], (uint32_t)(52 - i)));
                    free_mask &= ~bit;
                    if (i < 5) board |= bit;  // draws 0..4: the board (game.py:278)
                    else if (i & 1) pair = bit;                          // draw 5 + 2p: first hole card of seat p
                    else pd[(size_t)((i - 6) >> 1) * ld + s] = pair | bit;  // draw 6 + 2p: second hole card of seat p
                }
            }
        }
    }
    reinterpret_cast<uint64_t*>(v.mr)[s] = board;
}

// :488-489 eval_hand(board + hole) for one seat that shows its cards
template <class V>
PKB_D void showdown_rank(V& v, size_t s, int seat)
{
    uint64_t* word = reinterpret_cast<uint64_t*>(v.pd) + seat * v.ld() + s;
    *word = eval_mask52_sel(reinterpret_cast<const uint64_t*>(v.mr)[s] | *word);
}

template <int P, class V>
PKB_D void showdown_settle(V& v, size_t s, const Seats& x, int npw)
{
    const size_t ld = v.ld();
    uint32_t rk[P];
    double b[P], pay[P];
    uint32_t todo = x.called | x.allin;  // :488, :496 seats that show their cards; NONE for the others
#pragma unroll
    for (int p = 0; p < P; ++p) rk[p] = (todo >> p & 1u) ? (uint32_t)reinterpret_cast<const uint64_t*>(v.pd)[p * ld + s] : PACKED_NONE;
    double bet0[P];  // self.bets: the order of the layers (:496) goes by the bets as they were
#pragma unroll
    for (int p = 0; p < P; ++p) {
        b[p] = bet0[p] = v.bt[p * ld + s];  // :485 working copy
        pay[p] = 0.0;
    }
    int n = npw;
    while (todo) {  // :496-498 seats by ascending bet, ties by seat (np.argsort, numpy 1.18.4 order)
        int p = 0;
        double least = 0.0;
        bool have = false;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const double bi = bet0[i];
            if ((todo >> i & 1u) && (!have || bi < least)) { p = i; least = bi; have = true; }
        }
        todo &= ~(1u << p);
        bool nothing_left = true;
#pragma unroll
        for (int i = 0; i < P; ++i) nothing_left &= b[i] <= 0.0;
        if (nothing_left) break;  // :499
        if (n == 1) {             // :500-505
            const double rest = np_sum<P>(b);
#pragma unroll
            for (int i = 0; i < P; ++i)
                if (i == p) pay[i] += rest;
            break;
        }
        double max_bet = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i)
            if (i == p) max_bet = b[i];  // :508
        double layer[P];
#pragma unroll
        for (int i = 0; i < P; ++i) {  // :509 np.clip(bets, 0, max_bet)
            const double bi = b[i] < 0.0 ? 0.0 : b[i];
            layer[i] = bi > max_bet ? max_bet : bi;
        }
        const uint32_t winners = winners_closed<P>(rk);  // :512 compare_rankings
        const double S = np_sum<P>(layer);
        const int nw = popc32(winners);
        // :515 one winner takes S; :516 otherwise payoffs += (S * onehot) / sum(onehot).  A winner's share is S * 1.0 / nw;
        // everybody else's is (S * 0.0) / nw, a signed zero, and adding a zero of either sign leaves a payoff as it is
        // (payoffs start at +0.0 and a sum is -0.0 only if both terms are): ONE division per layer, not one per seat.
        // Two seats can only tie two ways, and x / 2.0 == x * 0.5 exactly: no division sequence in the heads-up kernel.
        const double share = nw == 1 ? S : P == 2 ? __dmul_rn(S, 0.5) : S / (double)nw;
#pragma unroll
        for (int i = 0; i < P; ++i)
            if (winners >> i & 1u) pay[i] += share;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            if (i == p) rk[i] = PACKED_NONE;  // :522
            b[i] = b[i] - layer[i];           // :523
        }
        --n;  // :525
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        v.cr[p * ld + s] += pay[p];  // :528
        v.pd[p * ld + s] = pay[p];
    }
    v.mr[s] = 0.0;  // as end_hand's commit left it
}

// Game.end_hand, last part (game.py:530-539), then what the caller of end_hand does with its
// return value (game.py:565 / 609-611 or game.py:619).  `pay_out` receives the hand's final payoffs.
// With `auto_reset`, a finished game is reset right here (Game.reset, game.py:397-412).  Returns true if it reset.
template <int P, class V>
PKB_D bool end_hand_finish(V& v, size_t s, Seats& x, Meta& m, Ctl& c, double* pay_record, double (&pay_out)[P], const GameConst& gc)
{
    const size_t ld = v.ld();
    double rec[payoff_pitch(P)];
#pragma unroll
    for (int p = 0; p < payoff_pitch(P); ++p) rec[p] = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const double pay = v.pd[p * ld + s] - v.bt[p * ld + s];  // :531
        pay_out[p] = rec[p] = pay;
        if (v.cr[p * ld + s] <= 0.0) set_state(x, p, ST_BROKEN);  // :536
    }
#ifndef PKB_PROBE_NO_PAYOFF_STORE  // timing probe only (profiles/): what the payoff stores cost
#pragma unroll
    for (int q = 0; q < payoff_pitch(P); q += 4) store_f64x4(pay_record + q, rec[q], rec[q + 1], rec[q + 2], rec[q + 3]);
#endif
    // :539 setup_hand runs even when the game is over -- and it can bring BROKEN seats with a negative
    // stack back as ALL_IN, so game_over is read after it, as the callers do
    // (:565, :619).
    setup_hand<P>(v, s, x, m, gc);
    const bool over = game_over<P>(x);
    const bool reset_now = over && gc.auto_reset && !c.poisoned;
    if (reset_now) {  // Game.reset (game.py:397-412): its outcome depends on the configuration only
        v.hand[s] = 1;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            v.cr[p * ld + s] = gc.start_credits[p];
            v.pd[p * ld + s] = gc.fresh_pending[p];
        }
        v.mr[s] = gc.fresh_min_raise;
        m = unpack_meta(gc.fresh_meta);
        x = Seats{gc.fresh_active, gc.fresh_called, gc.fresh_allin, 0};
    }
    const int flag = over ? (int)F_GAME_OVER : 0;
    if (c.cont) {  // next_turn: `return self.game_over, True, True`, then the scan goes on (:609-611)
        c.done = flag | (int)(F_HAND_OVER | F_ROUND_OVER);
        if (over) c.pc = PC_DONE;
        else seat_scan<P>(v, s, x, m, c);
    } else {  // :619
        c.done = flag | (int)F_HAND_OVER;
        c.pc = PC_DONE;
    }
    return reset_now;
}

// ---- PokerGameEnv (envs/game_env.py:6-54) as a per-env automaton ---------------------------------------------
//
// PokerGameEnv.step plays the agent's action and then lets the other seats act until the hand ends or it is the agent's
// turn again -- two `while` loops around Game.step.  The env kernel runs those loops round by round for a whole tile:
// round 0 is the agent's Game.step, every further round one opponent move in the envs whose loops have not returned.
// env_advance is the control flow between two rounds of ONE env: given the control byte, the flag byte of the Game.step
// just made, the seat states and the new active player, it returns the next control byte.
//   bits 0-1 phase (0 returned / idle, 1 inside `while not hand and active != agent` :41-44, 2 inside `while not done and
//   active != agent` :49-52, 3 inside reset()'s `while active != agent` :24-27), bit 2 another opponent move is due,
//   bit 3 done, bit 4 hand_over, bit 5 the env was not stepped (PKB_ACTION_SKIP), bit 6 the agent's action was rejected
//   (ValueError in the reference), bit 7 poisoned (the reference would not terminate / would assert).
// Transient bits, never stored: AGENT_STEP (in: this was the agent's own step), WRITE_REWARD + REWARD (out: the reward is
// decided now -- payoffs[agent] if REWARD else 0.0 -- :39 and :47; it must be read NOW, the second loop may end more
// hands), REDO (out, reset loop: `if done: self.game.reset()` :27).
enum : uint32_t { EC_PHASE = 3u, EC_LOOP1 = 1u, EC_LOOP2 = 2u, EC_RESET_LOOP = 3u, EC_PENDING = 4u, EC_DONE = 8u, EC_HAND = 16u,
                  EC_SKIPPED = 32u, EC_REJECTED = 64u, EC_POISONED = 128u,
                  EC_AGENT_STEP = 0x100u, EC_WRITE_REWARD = 0x200u, EC_REWARD = 0x400u, EC_REDO = 0x800u };
constexpr int ENV_ROUND_CAP = 1024;  // opponent moves inside one PokerGameEnv call before the env is given up as poisoned

// the byte the caller gets: bit 0 done, bit 1 hand_over, bits 5-7 as above
PKB_D uint32_t env_out_byte(uint32_t ec) { return (ec >> 3 & 3u) | (ec & (EC_SKIPPED | EC_REJECTED | EC_POISONED)); }

template <int P>
PKB_D uint32_t env_advance(uint32_t ec_in, uint32_t flags, const Seats& x, int active, bool cap_hit)
{
    const bool done_now = (flags & F_GAME_OVER) != 0, hand_now = (flags & F_HAND_OVER) != 0, dead = (flags & F_POISONED) != 0;
    const uint32_t gave_up = EC_POISONED | EC_DONE | EC_HAND;  // treated as a finished game: the caller's reset re-arms the env
    const bool agents_turn = active == 0;                      // self.player_agent = 0 (game_env.py:18)
    uint32_t ec;
    int phase = (int)(ec_in & EC_PHASE);
    if (phase == (int)EC_RESET_LOOP) {  // :24-27
        if (done_now || dead) return EC_RESET_LOOP | EC_REDO;
        return EC_RESET_LOOP | (!agents_turn && !cap_hit ? EC_PENDING : 0u);
    }
    if (ec_in & EC_AGENT_STEP) {  // :34-39
        if (flags & F_SKIPPED) return EC_SKIPPED | EC_WRITE_REWARD;
        if (flags & F_INVALID) return EC_REJECTED | EC_WRITE_REWARD;
        if (dead) return gave_up | EC_WRITE_REWARD;
        if (done_now || (x.broken & 1u)) return EC_DONE | EC_HAND | EC_WRITE_REWARD | EC_REWARD;  // :37-39
        ec = (done_now ? EC_DONE : 0u) | (hand_now ? EC_HAND : 0u);
        phase = (int)EC_LOOP1;
    } else if (phase == (int)EC_LOOP1) {  // `done, hand, _ = self.game.step(action)` :44
        if (dead) return gave_up | EC_WRITE_REWARD;
        ec = (done_now ? EC_DONE : 0u) | (hand_now ? EC_HAND : 0u);
    } else {  // EC_LOOP2: `done, *_ = self.game.step(action)` :52 -- hand_over keeps the value of the first loop
        if (dead) return gave_up;
        ec = (ec_in & EC_HAND) | (done_now ? EC_DONE : 0u);
    }
    if (phase == (int)EC_LOOP1) {
        if (!(ec & EC_HAND) && !agents_turn) return cap_hit ? gave_up | EC_WRITE_REWARD : ec | EC_LOOP1 | EC_PENDING;  // :41
        ec |= EC_WRITE_REWARD | ((ec & EC_HAND) ? EC_REWARD : 0u);  // :47
    }
    if (!(ec & EC_DONE) && !agents_turn) return cap_hit ? ec | gave_up : ec | EC_LOOP2 | EC_PENDING;  // :49
    return ec;  // :54
}

}  // namespace pkb
