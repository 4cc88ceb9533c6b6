// pkb_game.cuh -- the reference Game state machine (pokerl/game.py) as device functions over a
// structure-of-arrays view of N lockstep environments.
//
// Everything here follows the reference line by line, including its non-standard behaviours
// (SURVEY.md Appendix C): money is IEEE double, products are never fused into adds
// (game.py:370-371, 675-678), sums follow numpy's add.reduce order, and the winner scan keeps
// the judger.py:148 behaviour.  A View addresses element (seat p, slot s) of a plane at
// plane[p * ld + s]; the same code runs on a shared-memory tile (ld = tile width) and directly
// on the HBM planes (ld = padded env count).
#pragma once
#include "pkb_common.cuh"
#include "pkb_eval.cuh"
#include "pkb_philox.cuh"

namespace pkb {

constexpr uint32_t F_GAME_OVER = 1u, F_HAND_OVER = 2u, F_ROUND_OVER = 4u, F_INVALID = 64u, F_POISONED = 128u;
constexpr int NEXT_TURN_CAP = 64;
constexpr int MAX_PLAYERS = 10;

// per-handle constants, passed by value to every kernel (Game(**config), game.py:246-251)
struct GameConst {
    double start_credits[MAX_PLAYERS];
    double big_blind, small_blind;
    uint64_t seed;
    uint32_t env_id_base;
    int32_t dealer;
    int32_t auto_reset;
    int32_t track_stats;
};

template <int LD>
struct View {
    double* cr;      // credits          [P][ld]
    double* bt;      // bets             [P][ld]
    double* pd;      // pending_bets     [P][ld]
    double* mr;      // minimum_raise    [ld]
    uint8_t* st;     // player_states    [P][ld]
    uint32_t* meta;  // packed indices   [ld]
    uint32_t* hand;  // hand counter     [ld]
    size_t ld_rt;
    PKB_D size_t ld() const { return LD ? (size_t)LD : ld_rt; }
};

// turn | dealer<<3 | small_blind_idx<<8 | big_blind_idx<<13 | active_player<<18
struct Meta {
    int turn, dealer, sb, bb, act;
};
PKB_D Meta unpack_meta(uint32_t w)
{
    return Meta{(int)(w & 7u), (int)(w >> 3 & 31u), (int)(w >> 8 & 31u), (int)(w >> 13 & 31u), (int)(w >> 18 & 31u)};
}
PKB_D uint32_t pack_meta(const Meta& m)
{
    return (uint32_t)m.turn | (uint32_t)m.dealer << 3 | (uint32_t)m.sb << 8 | (uint32_t)m.bb << 13 | (uint32_t)m.act << 18;
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
template <int P, class V>
PKB_D int first_playing(const V& v, size_t s, int idx)
{
    if (idx >= P) idx -= P;
    int found = -1;
#pragma unroll
    for (int k = 0; k < P; ++k) {
        int i = idx + k;
        if (i >= P) i -= P;
        if (found < 0 && v.st[i * v.ld() + s] != ST_BROKEN) found = i;
    }
    return found < 0 ? idx : found;
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
template <int P, class V>
PKB_D bool game_over(const V& v, size_t s)
{
    int n = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) n += v.st[p * v.ld() + s] != ST_BROKEN;
    return n == 1;
}

// Game.get_valid_actions, game.py:339-383 -> bit a set iff move a is valid
template <int P, class V>
PKB_D uint32_t valid_mask(const V& v, size_t s, int player)
{
    const double hb = high_bet<P>(v, s), credit = v.cr[player * v.ld() + s], mr = v.mr[s];
    const double room = credit - hb;
    uint32_t m = 1u << MV_FOLD | 1u << MV_ALL_IN;  // :367
    const double f[3] = {0.1, 0.25, 0.5};
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // :370-372, product and sum rounded separately
        const double rv = __dmul_rn(f[i], room);
        if (rv > mr && __dadd_rn(hb, rv) < credit) m |= 1u << (MV_RAISE + i);
    }
    if (hb == 0.0) m |= 1u << MV_CHECK;    // :375
    if (hb < credit) m |= 1u << MV_CALL;   // :376
    return m;
}

// Game.setup_hand, game.py:414-451.  The deck is not materialised: hand `hand` of this episode is
// a pure function of (seed, env, episode, hand) and is regenerated where it is needed.
template <int P, class V>
PKB_D void setup_hand(V& v, size_t s, Meta& m, const GameConst& gc)
{
    const size_t ld = v.ld();
    v.hand[s] += 1;
    m.turn = 0;
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (v.st[p * ld + s] != ST_BROKEN) v.st[p * ld + s] = ST_ACTIVE;  // :421
    m.dealer = first_playing<P>(v, s, m.dealer + 1);                      // :432
    m.sb = first_playing<P>(v, s, m.dealer + 1);                          // :433
    m.bb = first_playing<P>(v, s, m.sb + 1);                              // :434
    m.act = first_playing<P>(v, s, m.bb + 1);                             // :435
#pragma unroll
    for (int p = 0; p < P; ++p) {  // :438-440; the small blind is written last, so it wins when sb == bb
        v.bt[p * ld + s] = 0.0;
        v.pd[p * ld + s] = p == m.sb ? gc.small_blind : p == m.bb ? gc.big_blind : 0.0;
    }
    v.st[m.bb * ld + s] = ST_CALLED;  // :441
    double hb = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {  // :444-446
        double x = v.pd[p * ld + s];
        const double c = v.cr[p * ld + s];
        if (x > c) v.st[p * ld + s] = ST_ALL_IN;
        if (!(x <= c)) { x = c; v.pd[p * ld + s] = x; }
        hb = (p == 0 || x > hb) ? x : hb;
    }
    v.mr[s] = hb;
}

// Game.reset, game.py:397-412 (the episode counter is bumped by the caller)
template <int P, class V>
PKB_D void reset_env(V& v, size_t s, Meta& m, const GameConst& gc)
{
    const size_t ld = v.ld();
    m.dealer = gc.dealer;  // :403
    m.act = 0;
    v.hand[s] = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        v.cr[p * ld + s] = gc.start_credits[p];  // :408
        v.st[p * ld + s] = ST_ACTIVE;            // :409
    }
    setup_hand<P>(v, s, m, gc);
}

// game.py:457-461 and 554-558
template <int P, class V>
PKB_D void commit(V& v, size_t s)
{
    const size_t ld = v.ld();
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const double x = v.pd[p * ld + s];
        v.bt[p * ld + s] += x;
        v.cr[p * ld + s] -= x;
        v.pd[p * ld + s] = 0.0;
    }
    v.mr[s] = 0.0;
}

// The betting part of Game.step, game.py:656-696 (the action is already known to be valid)
template <int P, class V>
PKB_D void apply_action(V& v, size_t s, int p, int action, const GameConst& gc)
{
    const size_t ld = v.ld();
    if (action == MV_FOLD) v.st[p * ld + s] = ST_FOLDED;        // :656-657
    else if (action == MV_CHECK) v.st[p * ld + s] = ST_CALLED;  // :659-660
    else {
        const double hb = high_bet<P>(v, s);                     // :664
        double bet = hb >= gc.big_blind ? hb : gc.big_blind;     // :665 max(): first argument on ties
        const double credit = v.cr[p * ld + s];
        uint8_t mine = ST_CALLED;                                // :667
        if (action == MV_ALL_IN) {                               // :669-671
            bet = credit;
            mine = ST_ALL_IN;
        } else if (action >= MV_RAISE) {                         // :673-678, three rounded operations
            const double f = action == MV_RAISE ? 0.1 : action == MV_RAISE + 1 ? 0.25 : 0.5;
            bet = __dadd_rn(bet, __dmul_rn(credit - bet, f));
        }
        if (bet > hb) {  // :680-687
#pragma unroll
            for (int i = 0; i < P; ++i)
                if (v.st[i * ld + s] == ST_CALLED) v.st[i * ld + s] = ST_ACTIVE;
            v.mr[s] = bet - hb;
        }
        v.st[p * ld + s] = mine;
        v.pd[p * ld + s] = bet;  // :696
    }
}

// ---- the resumable part of Game.next_player / next_turn / end_hand ------------------------------
//
// A step is run as a small automaton so that the expensive, rarely taken branches (end_hand and
// the showdown inside it) can be executed by densely packed threads instead of diverging warps.
// Control word, one per env slot:
//   bits 0-1 pc (0 done, 1 needs end_hand, 2 needs showdown), bit 2 cont (1: end_hand was entered
//   from next_turn and the seat scan resumes afterwards, game.py:609-611; 0: from the fold path,
//   game.py:616-619), bits 3-7 current_player of the scan, bits 8-10 the `done` tuple so far,
//   bits 11-17 next_turn() calls made, bit 18 poisoned, bit 19 invalid action,
//   bits 20-23 number of potential winners (showdown hand-over between phases)
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

// The `while` of Game.next_player (game.py:607-611) with next_turn (game.py:541-576) inlined, up to
// the point where next_turn would call end_hand (turn == 4).
template <int P, class V>
PKB_D void seat_scan(V& v, size_t s, Meta& m, Ctl& c)
{
    const size_t ld = v.ld();
    c.pc = PC_DONE;
    while (v.st[m.act * ld + s] != ST_ACTIVE) {
        if (m.act == c.cur) {
            if (c.nt >= NEXT_TURN_CAP) {  // the reference has no termination guarantee here
                c.poisoned = 1;
                return;
            }
            ++c.nt;
            commit<P>(v, s);  // :554-558
            m.turn += 1;      // :561
            if (m.turn == 4) {  // :563-565 -> end_hand, then `return game_over, True, True`
                c.pc = PC_END_HAND;
                c.cont = 1;
                return;
            }
            int called = 0;
#pragma unroll
            for (int p = 0; p < P; ++p) called += v.st[p * ld + s] == ST_CALLED;
            if (called > 1) {  // :570-572
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if (v.st[p * ld + s] == ST_CALLED) v.st[p * ld + s] = ST_ACTIVE;
            }
            m.act = first_playing<P>(v, s, m.dealer + 1);  // :575
            c.done = (int)F_ROUND_OVER;                    // :576
        } else {
            m.act = m.act + 1 == P ? 0 : m.act + 1;  // :611
        }
    }
}

// Game.step up to the first end_hand (game.py:648-699 and 597-605)
template <int P, class V>
PKB_D void step_begin(V& v, size_t s, Meta& m, Ctl& c, int action, uint32_t vmask, const GameConst& gc)
{
    const size_t ld = v.ld();
    c = Ctl{PC_DONE, 0, 0, 0, 0, 0, 0, 0};
    if (action < 0 || action >= NUM_MOVES || !(vmask >> action & 1u)) {  // :648-651, before any mutation
        c.invalid = 1;
        return;
    }
    apply_action<P>(v, s, m.act, action, gc);
    int playing = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int st = v.st[p * ld + s];
        playing += st != ST_BROKEN && st != ST_FOLDED;  // :598-599
    }
    if (playing > 1) {  // :601-605
        c.cur = m.act;
        m.act = m.act + 1 == P ? 0 : m.act + 1;
        seat_scan<P>(v, s, m, c);
    } else {  // :616-619
        c.pc = PC_END_HAND;
        c.cont = 0;
    }
}

// Game.end_hand, first part (game.py:453-482): commit, count the potential winners, settle a
// last-stand pot.  Returns true when the hand needs the ranked showdown (game.py:483-528).
// From here until setup_hand the PENDING plane carries self.payoffs (it is all zero after commit and
// is rewritten by the next setup_hand), so no per-seat register array crosses a phase boundary.
template <int P, class V>
PKB_D bool end_hand_begin(V& v, size_t s, Ctl& c)
{
    const size_t ld = v.ld();
    commit<P>(v, s);  // :457-461; pending (= payoffs from here on) is now 0.0, :468
    int n = 0, first = -1;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int st = v.st[p * ld + s];
        if (st != ST_BROKEN && st != ST_FOLDED) {  // :471
            if (first < 0) first = p;
            ++n;
        }
    }
    c.npw = n;
    if (n == 0) {  // :473 -- AssertionError in the reference
        c.poisoned = 1;
        return false;
    }
    if (n == 1) {  // :475-482
        const double pot = sum_plane<P>(v, v.bt, s);
        v.pd[first * ld + s] = pot;
        v.cr[first * ld + s] += pot;
        return false;
    }
    return true;
}

// Game.end_hand, ranked showdown (game.py:483-528).  `rank7(mask)` evaluates a 7-card id mask
// (eval_mask52, possibly out of line).  The working copy of bets is kept in registers; payoffs
// are written to the pending plane at the end.
template <int P, class V, class Rank7>
PKB_D void showdown(V& v, size_t s, const Ctl& c, uint64_t seed, uint32_t env_id, uint32_t episode, Rank7 rank7)
{
    const size_t ld = v.ld();
    uint32_t rk[P];
    double b[P], pay[P];
    uint32_t todo = 0;
    {  // :488-489 eval_hand(board + hole) for CALLED / ALL_IN seats, NONE for the others
        Dealer d(seed, env_id, episode, v.hand[s]);
        uint64_t board = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) board |= 1ull << d.next();
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const uint64_t hole = (1ull << d.next()) | (1ull << d.next());
            const int st = v.st[p * ld + s];
            const bool shows = st == ST_CALLED || st == ST_ALL_IN;
            rk[p] = shows ? rank7(board | hole) : PACKED_NONE;
            todo |= (uint32_t)shows << p;
            b[p] = v.bt[p * ld + s];  // :485 working copy
            pay[p] = 0.0;
        }
    }
    int n = c.npw;
    while (todo) {  // :496-498 seats by ascending bet, ties by seat (np.argsort, numpy 1.18.4 order)
        int p = 0;
        double least = 0.0;
        bool have = false;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const double x = v.bt[i * ld + s];
            if ((todo >> i & 1u) && (!have || x < least)) { p = i; least = x; have = true; }
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
            const double x = b[i] < 0.0 ? 0.0 : b[i];
            layer[i] = x > max_bet ? max_bet : x;
        }
        WinnerScan ws;  // :512
#pragma unroll
        for (int i = 0; i < P; ++i) ws.feed(i, rk[i]);
        const double S = np_sum<P>(layer);
        const int nw = popc32(ws.winners);
        if (nw == 1) {  // :515
#pragma unroll
            for (int i = 0; i < P; ++i)
                if (ws.winners >> i & 1u) pay[i] += S;
        } else {  // :516 (S * onehot) / sum(onehot)
            const double k = (double)nw;
#pragma unroll
            for (int i = 0; i < P; ++i) pay[i] += __dmul_rn(S, (ws.winners >> i & 1u) ? 1.0 : 0.0) / k;
        }
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
}

// Game.end_hand, last part (game.py:530-539), then what the caller of end_hand does with its
// return value (game.py:565 / 609-611 or game.py:619).  `deal(m)` runs setup_hand (possibly out of
// line) and returns the new Meta; `pay_out` receives the hand's final payoffs.
template <int P, class V, class Deal>
PKB_D void end_hand_finish(V& v, size_t s, Meta& m, Ctl& c, double* payoffs_out, size_t out_ld, size_t out_idx,
                           double (&pay_out)[P], Deal deal)
{
    const size_t ld = v.ld();
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const double pay = v.pd[p * ld + s] - v.bt[p * ld + s];  // :531
        pay_out[p] = pay;
        payoffs_out[p * out_ld + out_idx] = pay;
        if (v.cr[p * ld + s] <= 0.0) v.st[p * ld + s] = ST_BROKEN;  // :536
    }
    m = deal(m);  // :539 setup_hand -- runs even when the game is over
    const int over = game_over<P>(v, s) ? (int)F_GAME_OVER : 0;
    if (c.cont) {  // next_turn: `return self.game_over, True, True`, then the scan goes on (:609-611)
        c.done = over | (int)(F_HAND_OVER | F_ROUND_OVER);
        if (over) c.pc = PC_DONE;
        else seat_scan<P>(v, s, m, c);
    } else {  // :619
        c.done = over | (int)F_HAND_OVER;
        c.pc = PC_DONE;
    }
}

}  // namespace pkb
