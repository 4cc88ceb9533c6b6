/*
 * pokerl_oracle.c -- CPU oracle for the pokerl hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Scalar restatement of sneppy/pokerl v0.0.6 (judger.py, game.py, cards.py).
 * Every function cites the reference lines it follows.  Used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the CHECKER / CPU comparator; never by the product path.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp)
 * Floating point: IEEE double, no contraction -- the reference computes
 * a*b+c as two rounded numpy/Python operations (game.py:370-371, 675-678).
 */
#include "pokerl_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ cards */

static inline int card_rank(uint8_t v) /* cards.py:8-14 */
{
    int r = v & 0xf;
    return r ? r : 13;
}
static inline int card_suit(uint8_t v) { return v >> 4; } /* cards.py:17-20 */

static inline uint8_t card_from_id(int id) /* inverse of cards.py:22-26 */
{
    return (uint8_t)(((id / 13) << 4) | (id % 13));
}

/* ----------------------------------------------------------------- judger */

/* get_kickers_value, judger.py:101-109: first kicker is the most significant nibble */
static uint32_t kickers_value(const int *k, int n)
{
    uint32_t v = 0;
    for (int i = 0; i < n; ++i) v |= (uint32_t)k[n - 1 - i] << (4 * i);
    return v;
}

#define PACK(cat, k, n) (((uint32_t)(cat) << 20) | kickers_value((k), (n)))

/* stable descending sort by key (Python sorted(..., reverse=True) keeps the
 * original order of equal keys), judger.py:38-39 */
static void sort_desc(const uint8_t *in, int n, uint8_t *out, int by_suit)
{
    int key[64];
    for (int i = 0; i < n; ++i) {
        int k = by_suit ? ((card_suit(in[i]) << 4) | card_rank(in[i])) : card_rank(in[i]);
        int j = i;
        while (j > 0 && key[j - 1] < k) {
            key[j] = key[j - 1];
            out[j] = out[j - 1];
            --j;
        }
        key[j] = k;
        out[j] = in[i];
    }
}

/* eval_hand, judger.py:7-99.  Literal: two sorted copies scanned in lock step. */
uint32_t ora_eval_cards(const uint8_t *hand, int n)
{
    int kk[8];
    if (n > 64) n = 64;
    if (n <= 0) return 10u << 20; /* :30 */
    if (n == 1) {                 /* :31 */
        kk[0] = card_rank(hand[0]);
        return PACK(9, kk, 1);
    }
    if (n == 2) { /* :32-35 */
        int a = card_rank(hand[0]), b = card_rank(hand[1]);
        if (a == b) {
            kk[0] = a;
            return PACK(8, kk, 1);
        }
        kk[0] = a > b ? a : b;
        kk[1] = a > b ? b : a;
        return PACK(9, kk, 2);
    }

    uint8_t rs[64], ss[64];
    sort_desc(hand, n, rs, 0); /* :38 */
    sort_desc(hand, n, ss, 1); /* :39 */

    int flush = 4, both = 4, kind = 0, straight = 0, flush_start = 0; /* :41-45 */
    int four[16], three[21], two[32];
    int nfour = 0, nthree = 0, ntwo = 0;

    for (int idx = 0; idx < n; ++idx) { /* :50 */
        int srank = card_rank(ss[idx]), ssuit = card_suit(ss[idx]);
        if (ssuit == (flush & 0xf)) { /* :52 */
            flush += 0x100;
            if (srank + (both >> 8) == ((both >> 4) & 0xf)) both += 0x100; /* :56 */
            else both = 0x100 | (srank << 4) | ssuit;                     /* :57 */
        } else if ((flush >> 8) < 5) {                                    /* :58 */
            both = flush = 0x100 | (srank << 4) | ssuit;
            flush_start = idx;
        }

        int rr = card_rank(rs[idx]);
        if (rr == (kind & 0xf)) kind += 0x10; /* :61 */
        else {
            int num = kind >> 4; /* :64-67 */
            if (num == 2) two[ntwo++] = kind & 0xf;
            else if (num == 3) three[nthree++] = kind & 0xf;
            else if (num == 4) four[nfour++] = kind & 0xf;
            kind = 0x10 | rr;                                                       /* :70 */
            if (rr + (straight >> 4) == (straight & 0xf)) straight += 0x10;         /* :71 */
            else if ((straight >> 4) < 5) straight = 0x10 | rr;                     /* :72 */
        }
    }
    { /* :77-80 */
        int num = kind >> 4;
        if (num == 2) two[ntwo++] = kind & 0xf;
        else if (num == 3) three[nthree++] = kind & 0xf;
        else if (num == 4) four[nfour++] = kind & 0xf;
    }

    /* five-high straight special cases, :83-88 */
    if ((both >> 8) == 4 && ((both >> 4) & 0xf) == 4) {
        for (int i = 0; i < n; ++i)
            if (card_rank(hand[i]) == 13 && card_suit(hand[i]) == (both & 0xf)) {
                kk[0] = 4;
                return PACK(1, kk, 1);
            }
    } else if ((straight >> 4) == 4 && (straight & 0xf) == 4) {
        for (int i = 0; i < n; ++i)
            if (card_rank(hand[i]) == 13) {
                kk[0] = 4;
                return PACK(5, kk, 1);
            }
    }

    int nk = 0;
    if ((both >> 8) >= 5) { /* :90 */
        kk[0] = (both >> 4) & 0xf;
        return PACK(1, kk, 1);
    } else if (nfour) { /* :91 */
        kk[nk++] = four[0];
        for (int i = 0; i < n && nk < 2; ++i)
            if (card_rank(rs[i]) != four[0]) kk[nk++] = card_rank(rs[i]);
        return PACK(2, kk, nk);
    } else if (nthree > 1) { /* :92 */
        kk[0] = three[0];
        kk[1] = three[1];
        return PACK(3, kk, 2);
    } else if (nthree && ntwo) { /* :93 */
        kk[0] = three[0];
        kk[1] = two[0];
        return PACK(3, kk, 2);
    } else if ((flush >> 8) >= 5) { /* :94 */
        for (int i = flush_start; i < n && nk < 5; ++i) kk[nk++] = card_rank(ss[i]);
        return PACK(4, kk, nk);
    } else if ((straight >> 4) >= 5) { /* :95 */
        kk[0] = straight & 0xf;
        return PACK(5, kk, 1);
    } else if (nthree) { /* :96 */
        kk[nk++] = three[0];
        for (int i = 0; i < n && nk < 3; ++i)
            if (card_rank(rs[i]) != three[0]) kk[nk++] = card_rank(rs[i]);
        return PACK(6, kk, nk);
    } else if (ntwo > 1) { /* :97 */
        kk[nk++] = two[0];
        kk[nk++] = two[1];
        for (int i = 0; i < n && nk < 3; ++i)
            if (card_rank(rs[i]) != two[0] && card_rank(rs[i]) != two[1]) kk[nk++] = card_rank(rs[i]);
        return PACK(7, kk, nk);
    } else if (ntwo) { /* :98 */
        kk[nk++] = two[0];
        for (int i = 0; i < n && nk < 4; ++i)
            if (card_rank(rs[i]) != two[0]) kk[nk++] = card_rank(rs[i]);
        return PACK(8, kk, nk);
    }
    for (int i = 0; i < n && nk < 5; ++i) kk[nk++] = card_rank(rs[i]); /* :99 */
    return PACK(9, kk, nk);
}

void ora_eval_cards_batch(const uint8_t *cards, int k, size_t n, uint32_t *out)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) out[i] = ora_eval_cards(cards + (size_t)i * k, k);
}

uint32_t ora_eval_mask(uint64_t mask)
{
    uint8_t cards[64];
    int n = 0;
    for (int id = 0; id < 52; ++id)
        if (mask >> id & 1) cards[n++] = card_from_id(id);
    return ora_eval_cards(cards, n);
}

void ora_eval_masks_batch(const uint64_t *masks, size_t n, uint32_t *out)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) out[i] = ora_eval_mask(masks[i]);
}

static uint64_t binom(int n, int k)
{
    if (k < 0 || k > n) return 0;
    uint64_t r = 1;
    for (int i = 1; i <= k; ++i) r = r * (uint64_t)(n - k + i) / (uint64_t)i;
    return r;
}

/* combinatorial number system: index -> k-subset mask, colex order */
static uint64_t unrank_colex(uint64_t m, int k)
{
    uint64_t mask = 0;
    for (int i = k; i >= 1; --i) {
        int c = i - 1;
        while (binom(c + 1, i) <= m) ++c;
        mask |= 1ull << c;
        m -= binom(c, i);
    }
    return mask;
}

static inline uint64_t next_same_popcount(uint64_t v) /* Gosper */
{
    uint64_t c = v & -v, r = v + c;
    return (((r ^ v) >> 2) / c) | r;
}

void ora_eval_enumerate(int k, uint64_t first, uint64_t count, uint32_t *out)
{
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    uint64_t chunk = (count + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t) {
        uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk < count ? lo + chunk : count;
        if (lo >= hi) continue;
        uint64_t mask = unrank_colex(first + lo, k);
        for (uint64_t i = lo; i < hi; ++i) {
            out[i] = ora_eval_mask(mask);
            mask = next_same_popcount(mask);
        }
    }
}

void ora_eval_exhaustive(int k, uint64_t hist[11], uint64_t *sumA, uint64_t *sumB)
{
    uint64_t total = binom(52, k);
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    uint64_t chunk = (total + nt - 1) / nt;
    uint64_t h[11] = {0}, a = 0, b = 0;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t) {
        uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk < total ? lo + chunk : total;
        if (lo >= hi) continue;
        uint64_t lh[11] = {0}, la = 0, lb = 0;
        uint64_t mask = unrank_colex(lo, k);
        for (uint64_t i = lo; i < hi; ++i) {
            uint32_t p = ora_eval_mask(mask);
            lh[p >> 20]++;
            la += p;
            lb += (uint64_t)p * (mask % 1000003ull);
            mask = next_same_popcount(mask);
        }
#pragma omp critical
        {
            for (int c = 0; c < 11; ++c) h[c] += lh[c];
            a += la;
            b += lb;
        }
    }
    memcpy(hist, h, sizeof h);
    *sumA = a;
    *sumB = b;
}

/* compare_rankings, judger.py:133-158 -- including line 148, which assigns
 * `kicker = best_kicker` instead of updating best_kicker. */
int ora_compare_rankings(const uint32_t *packed, int n, uint8_t *onehot)
{
    int winners[64], nw = 0;
    uint32_t best_rank = 10, best_kicker = 0;
    for (int idx = 0; idx < n; ++idx) {
        uint32_t rank = packed[idx] >> 20, kicker = packed[idx] & 0xfffff;
        if (rank < best_rank) { /* :140-144 */
            best_rank = rank;
            best_kicker = kicker;
            nw = 0;
            winners[nw++] = idx;
        } else if (rank == best_rank) {
            if (kicker > best_kicker) { /* :146-149, best_kicker unchanged */
                nw = 0;
                winners[nw++] = idx;
            } else if (kicker == best_kicker) /* :150-152 */
                winners[nw++] = idx;
        }
    }
    for (int i = 0; i < n; ++i) onehot[i] = 0;
    for (int i = 0; i < nw; ++i) onehot[winners[i]] = 1;
    return nw;
}

/* ------------------------------------------------------- Philox / dealing */

/* Philox4x32-10 (Salmon et al., SC'11), this repo's dealing/sampling RNG. */
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/*
 * Deal spec (DESIGN.md "Dealing"): draw i of a hand (i = 0 .. 4+2P) uses 32-bit
 * word (i & 3) of Philox block (i >> 2) with counter
 *   {env_id, hand, episode, block} and key {seed_lo, seed_hi};
 * j = mulhi32(word, 52 - i) picks the j-th smallest card id not drawn yet.
 * Replaces random.shuffle(self.deck) at game.py:424; positions follow
 * game.py:278, 385-395 (board = deck[:5], hole p = deck[5+2p:7+2p]).
 */
void ora_deal(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t hand, int num_players, uint8_t *cards)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint64_t freemask = (1ull << 52) - 1;
    uint32_t blk[4];
    int total = 5 + 2 * num_players;
    for (int i = 0; i < total; ++i) {
        if ((i & 3) == 0) {
            uint32_t ctr[4] = {env_id, hand, episode, (uint32_t)(i >> 2)};
            ora_philox4x32_10(ctr, key, blk);
        }
        uint32_t j = (uint32_t)(((uint64_t)blk[i & 3] * (uint64_t)(52 - i)) >> 32);
        int id = 0;
        for (;; ++id)
            if (freemask >> id & 1) {
                if (j == 0) break;
                --j;
            }
        freemask &= ~(1ull << id);
        cards[i] = card_from_id(id);
    }
}

/*
 * Action sampling spec: counter {env_id & ~96, step_idx, 0, 0x80000000}; word (env_id >> 5) & 3
 * (the four envs whose ids differ only in bits 5 and 6 share one Philox block, one word each);
 * t = mulhi32(word, W) where W = sum of weights of valid actions; the chosen
 * action is the first valid a with cumulative weight > t.  Uniform = all 1
 * (RandomAgent, agents/random.py:16-17, is uniform over the valid actions).
 */
int ora_sample_action(uint64_t seed, uint32_t env_id, uint32_t step_idx, unsigned valid_mask, const uint32_t *weights)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {env_id & ~96u, step_idx, 0u, 0x80000000u}, out[4];
    ora_philox4x32_10(ctr, key, out);
    const uint32_t word = out[env_id >> 5 & 3u];
    uint32_t total = 0;
    for (int a = 0; a < ORA_NUM_MOVES; ++a)
        if (valid_mask >> a & 1) total += weights ? weights[a] : 1u;
    uint32_t t = (uint32_t)(((uint64_t)word * (uint64_t)total) >> 32), acc = 0;
    for (int a = 0; a < ORA_NUM_MOVES; ++a)
        if (valid_mask >> a & 1) {
            acc += weights ? weights[a] : 1u;
            if (acc > t) return a;
        }
    return ORA_FOLD; /* unreachable when total > 0: FOLD is always valid */
}

/* ------------------------------------------------------------------ numpy */

/* np.sum of a contiguous float64 vector (numpy pairwise_sum) */
double ora_np_sum(const double *a, int n)
{
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
    int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i + 8 <= n; i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

/* ------------------------------------------------------------------- game */

void ora_game_init(ora_game *g, int num_players, const double *start_credits, double big_blind, double small_blind,
                   int dealer, uint64_t seed, uint32_t env_id)
{
    memset(g, 0, sizeof *g); /* game.py:242-264 */
    g->num_players = num_players;
    for (int p = 0; p < num_players; ++p) {
        g->start_credits[p] = start_credits[p];
        g->states[p] = ORA_ACTIVE;
    }
    g->big_blind = big_blind;
    g->small_blind = small_blind;
    g->cfg_dealer = dealer;
    g->dealer = dealer;
    g->seed = seed;
    g->env_id = env_id;
}

void ora_game_set_decks(ora_game *g, const uint8_t *decks, int stride, int count)
{
    g->deck_queue = decks;
    g->deck_stride = stride;
    g->deck_count = count;
    g->deck_cursor = 0;
}

static int first_playing(const ora_game *g, int idx) /* game.py:334-337 */
{
    int P = g->num_players;
    for (int k = 0; k < P; ++k)
        if (g->states[(idx + k) % P] != ORA_BROKEN) return (idx + k) % P;
    return idx % P; /* argmax of all-False is 0 */
}

static double high_bet(const ora_game *g) /* game.py:286-290 */
{
    double m = g->pending[0];
    for (int p = 1; p < g->num_players; ++p)
        if (g->pending[p] > m) m = g->pending[p];
    return m;
}

int ora_game_is_over(const ora_game *g) /* game.py:316-320 */
{
    int n = 0;
    for (int p = 0; p < g->num_players; ++p) n += g->states[p] != ORA_BROKEN;
    return n == 1;
}

unsigned ora_game_valid_mask(const ora_game *g, int player) /* game.py:339-383 */
{
    static const double factors[3] = {0.1, 0.25, 0.5};
    if (player < 0) player = g->active_player;
    double hb = high_bet(g), credit = g->credits[player];
    unsigned m = 1u << ORA_FOLD | 1u << ORA_MOVE_ALL_IN; /* :367 */
    for (int i = 0; i < 3; ++i) {                        /* :370-372 */
        double rv = factors[i] * (credit - hb);
        if (rv > g->minimum_raise_value && (hb + rv) < credit) m |= 1u << (ORA_RAISE_TEN + i);
    }
    if (hb == 0.0) m |= 1u << ORA_CHECK;   /* :375 */
    if (hb < credit) m |= 1u << ORA_CALL;  /* :376 */
    return m;
}

static void setup_hand(ora_game *g) /* game.py:414-451 */
{
    int P = g->num_players;
    g->hand += 1;
    g->turn = 0;
    for (int p = 0; p < P; ++p)
        if (g->states[p] != ORA_BROKEN) g->states[p] = ORA_ACTIVE; /* :421 */

    /* :424 -- one deck consumed per setup_hand */
    if (g->deck_queue && g->deck_cursor < g->deck_count) {
        memcpy(g->deck, g->deck_queue + (size_t)g->deck_cursor * g->deck_stride, 5 + 2 * P);
        g->deck_cursor++;
    } else
        ora_deal(g->seed, g->env_id, g->episode, (uint32_t)g->hand, P, g->deck);

    g->dealer = first_playing(g, g->dealer + 1);                   /* :432 */
    g->small_blind_idx = first_playing(g, g->dealer + 1);          /* :433 */
    g->big_blind_idx = first_playing(g, g->small_blind_idx + 1);   /* :434 */
    g->active_player = first_playing(g, g->big_blind_idx + 1);     /* :435 */

    for (int p = 0; p < P; ++p) g->bets[p] = g->pending[p] = 0.0;  /* :438-439 */
    g->pending[g->big_blind_idx] = g->big_blind;                   /* :440, fancy index [bb, sb]: sb written last */
    g->pending[g->small_blind_idx] = g->small_blind;
    g->states[g->big_blind_idx] = ORA_CALLED;                      /* :441 */
    for (int p = 0; p < P; ++p) {
        if (g->pending[p] > g->credits[p]) g->states[p] = ORA_ALL_IN;             /* :444 */
        if (!(g->pending[p] <= g->credits[p])) g->pending[p] = g->credits[p];     /* :445 np.minimum */
    }
    g->minimum_raise_value = high_bet(g); /* :446 */
}

void ora_game_reset(ora_game *g) /* game.py:397-412 */
{
    g->dealer = g->cfg_dealer; /* :403 config.get('dealer', 0) of reset()'s OWN kwargs */
    g->hand = 0;
    g->active_player = 0;
    g->episode += 1;
    for (int p = 0; p < g->num_players; ++p) {
        g->credits[p] = g->start_credits[p];
        g->states[p] = ORA_ACTIVE;
    }
    setup_hand(g);
}

static void commit(ora_game *g) /* game.py:457-461 and 554-558 */
{
    for (int p = 0; p < g->num_players; ++p) {
        g->bets[p] += g->pending[p];
        g->credits[p] -= g->pending[p];
        g->pending[p] = 0.0;
    }
    g->minimum_raise_value = 0.0;
}

static void end_hand(ora_game *g) /* game.py:453-539 */
{
    int P = g->num_players;
    g->end_hand_calls++;
    commit(g);
    for (int p = 0; p < P; ++p) g->payoffs[p] = 0.0; /* :468 */

    int n = 0, first = -1;
    for (int p = 0; p < P; ++p)
        if (g->states[p] != ORA_BROKEN && g->states[p] != ORA_FOLDED) { /* :471 */
            if (first < 0) first = p;
            ++n;
        }
    if (n == 0) { /* :473 assert -> AssertionError in the reference */
        g->poisoned = 1;
        return;
    }

    if (n == 1) { /* :475-482 */
        double pot = ora_np_sum(g->bets, P);
        g->payoffs[first] = pot;
        g->credits[first] += pot;
    } else {
        double b[ORA_MAX_PLAYERS], layer[ORA_MAX_PLAYERS];
        uint32_t rk[ORA_MAX_PLAYERS];
        uint8_t onehot[ORA_MAX_PLAYERS];
        int order[ORA_MAX_PLAYERS];
        for (int p = 0; p < P; ++p) { /* :485-489 */
            b[p] = g->bets[p];
            if (g->states[p] == ORA_CALLED || g->states[p] == ORA_ALL_IN) {
                uint8_t hand7[7];
                memcpy(hand7, g->deck, 5); /* :395 board first */
                hand7[5] = g->deck[5 + 2 * p];
                hand7[6] = g->deck[6 + 2 * p];
                rk[p] = ora_eval_cards(hand7, 7);
            } else
                rk[p] = 10u << 20;
        }
        /* :495 np.argsort(bets): ascending, ties by seat (tie order does not change results) */
        for (int i = 0; i < P; ++i) {
            int j = i;
            while (j > 0 && b[order[j - 1]] > b[i]) {
                order[j] = order[j - 1];
                --j;
            }
            order[j] = i;
        }
        for (int oi = 0; oi < P; ++oi) { /* :496-525 */
            int p = order[oi];
            if (!(g->states[p] == ORA_CALLED || g->states[p] == ORA_ALL_IN)) continue;
            int all_nonpos = 1;
            for (int i = 0; i < P; ++i) all_nonpos &= b[i] <= 0.0;
            if (all_nonpos) break; /* :499 */
            if (n == 1) {          /* :500-505 */
                g->payoffs[p] += ora_np_sum(b, P);
                break;
            }
            double max_bet = b[p]; /* :508-509 np.clip(bets, 0, max_bet) = min(max(b,0), max_bet) */
            for (int i = 0; i < P; ++i) {
                double v = b[i] < 0.0 ? 0.0 : b[i];
                layer[i] = v > max_bet ? max_bet : v;
            }
            int nw = ora_compare_rankings(rk, P, onehot); /* :512 */
            double S = ora_np_sum(layer, P);
            g->showdown_layers++;
            if (nw == 1) { /* :515 */
                for (int i = 0; i < P; ++i)
                    if (onehot[i]) g->payoffs[i] += S;
            } else { /* :516 (S * onehot) / sum(onehot) */
                for (int i = 0; i < P; ++i) g->payoffs[i] += (S * (double)onehot[i]) / (double)nw;
            }
            rk[p] = 10u << 20;                            /* :522 */
            for (int i = 0; i < P; ++i) b[i] -= layer[i]; /* :523 */
            --n;                                          /* :525 */
        }
        for (int p = 0; p < P; ++p) g->credits[p] += g->payoffs[p]; /* :528 */
    }
    for (int p = 0; p < P; ++p) g->payoffs[p] -= g->bets[p]; /* :531 */
    for (int p = 0; p < P; ++p)
        if (g->credits[p] <= 0.0) g->states[p] = ORA_BROKEN; /* :536 */
    setup_hand(g);                                           /* :539 */
}

static unsigned next_turn(ora_game *g) /* game.py:541-576 */
{
    g->next_turn_calls++;
    commit(g);
    g->turn += 1;
    if (g->turn == 4) { /* :563-565 */
        end_hand(g);
        return (ora_game_is_over(g) ? ORA_F_GAME_OVER : 0) | ORA_F_HAND_OVER | ORA_F_ROUND_OVER;
    }
    int called = 0;
    for (int p = 0; p < g->num_players; ++p) called += g->states[p] == ORA_CALLED;
    if (called > 1) /* :570-572 */
        for (int p = 0; p < g->num_players; ++p)
            if (g->states[p] == ORA_CALLED) g->states[p] = ORA_ACTIVE;
    g->active_player = first_playing(g, g->dealer + 1); /* :575 */
    return ORA_F_ROUND_OVER;
}

static unsigned next_player(ora_game *g) /* game.py:578-619 */
{
    int P = g->num_players, playing = 0;
    for (int p = 0; p < P; ++p) playing += g->states[p] != ORA_BROKEN && g->states[p] != ORA_FOLDED;
    if (playing > 1) {
        unsigned done = 0;
        int current = g->active_player;
        g->active_player = (g->active_player + 1) % P;
        while (g->states[g->active_player] != ORA_ACTIVE) { /* :607 */
            if (current == g->active_player) {
                if (g->next_turn_calls >= ORA_NEXT_TURN_CAP) { /* no termination guarantee in the reference */
                    g->poisoned = 1;
                    return done | ORA_F_POISONED;
                }
                done = next_turn(g);
                if (g->poisoned) return done | ORA_F_POISONED;
                if (done & ORA_F_GAME_OVER) return done; /* :610 */
            } else
                g->active_player = (g->active_player + 1) % P; /* :611 */
        }
        return done;
    }
    end_hand(g); /* :616-619 */
    if (g->poisoned) return ORA_F_POISONED | ORA_F_HAND_OVER;
    return (ora_game_is_over(g) ? ORA_F_GAME_OVER : 0) | ORA_F_HAND_OVER;
}

unsigned ora_game_step(ora_game *g, int action) /* game.py:621-700 */
{
    static const double factors[3] = {0.1, 0.25, 0.5};
    g->next_turn_calls = g->end_hand_calls = g->showdown_layers = 0;
    g->poisoned = 0;
    if (action < 0 || action >= ORA_NUM_MOVES || !(ora_game_valid_mask(g, -1) >> action & 1))
        return ORA_F_INVALID; /* :648-651 ValueError before any mutation */

    int p = g->active_player, P = g->num_players;
    if (action == ORA_FOLD) g->states[p] = ORA_FOLDED;        /* :656-657 */
    else if (action == ORA_CHECK) g->states[p] = ORA_CALLED;  /* :659-660 */
    else {
        double hb = high_bet(g);                              /* :664 */
        double bet = hb >= g->big_blind ? hb : g->big_blind;  /* :665 max(high_bet, big_blind): first arg on ties */
        double credit = g->credits[p];
        g->states[p] = ORA_CALLED;                            /* :667 */
        if (action == ORA_MOVE_ALL_IN) {                      /* :669-671 */
            bet = credit;
            g->states[p] = ORA_ALL_IN;
        } else if (action >= ORA_RAISE_TEN) { /* :673-678, three separately rounded ops */
            double future = credit - bet;
            double raise = future * factors[action - ORA_RAISE_TEN];
            bet += raise;
        }
        if (bet > hb) { /* :680-687 */
            uint8_t mine = g->states[p];
            for (int i = 0; i < P; ++i)
                if (g->states[i] == ORA_CALLED) g->states[i] = ORA_ACTIVE;
            g->states[p] = mine;
            g->minimum_raise_value = bet - hb;
        }
        g->pending[p] = bet; /* :696 */
    }
    return next_player(g); /* :699 */
}

/* ------------------------------------------------------------------ batch */

struct ora_batch {
    size_t n;
    int auto_reset;
    ora_game *games;
};

static int g_threads = 0;
void ora_batch_set_threads(int nthreads)
{
    g_threads = nthreads;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
}

ora_batch *ora_batch_create(size_t n, int num_players, const double *start_credits, double big_blind, double small_blind,
                            int dealer, uint64_t seed, uint32_t env_id_base, int auto_reset)
{
    if (num_players < 1 || num_players > ORA_MAX_PLAYERS) return NULL;
    ora_batch *b = (ora_batch *)calloc(1, sizeof *b);
    b->n = n;
    b->auto_reset = auto_reset;
    b->games = (ora_game *)calloc(n ? n : 1, sizeof(ora_game));
    for (size_t i = 0; i < n; ++i)
        ora_game_init(&b->games[i], num_players, start_credits, big_blind, small_blind, dealer, seed,
                      env_id_base + (uint32_t)i);
    return b;
}

void ora_batch_destroy(ora_batch *b)
{
    if (!b) return;
    free(b->games);
    free(b);
}

ora_game *ora_batch_game(ora_batch *b, size_t i) { return &b->games[i]; }

void ora_batch_reset(ora_batch *b)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)b->n; ++i) ora_game_reset(&b->games[i]);
}

void ora_batch_valid(ora_batch *b, uint8_t *valid)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)b->n; ++i) valid[i] = (uint8_t)ora_game_valid_mask(&b->games[i], -1);
}

static inline unsigned step_one(ora_batch *b, ora_game *g, int action)
{
    unsigned f = ora_game_step(g, action);
    if ((f & ORA_F_GAME_OVER) && !(f & (ORA_F_INVALID | ORA_F_POISONED)) && b->auto_reset) ora_game_reset(g);
    return f;
}

void ora_batch_step(ora_batch *b, const int32_t *actions, uint8_t *flags, uint8_t *valid)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)b->n; ++i) {
        ora_game *g = &b->games[i];
        unsigned f = actions[i] == ORA_ACTION_SKIP ? ORA_F_SKIPPED : step_one(b, g, actions[i]);
        if (flags) flags[i] = (uint8_t)f;
        if (valid) valid[i] = (uint8_t)ora_game_valid_mask(g, -1);
    }
}

void ora_batch_sample(ora_batch *b, uint32_t step_idx, const uint32_t *weights, int32_t *actions)
{
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)b->n; ++i) {
        ora_game *g = &b->games[i];
        actions[i] = ora_sample_action(g->seed, g->env_id, step_idx, ora_game_valid_mask(g, -1), weights);
    }
}

uint64_t ora_batch_run_random(ora_batch *b, uint32_t step0, uint32_t steps, const uint32_t *weights)
{
    uint64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (long long i = 0; i < (long long)b->n; ++i) {
        ora_game *g = &b->games[i];
        for (uint32_t s = 0; s < steps; ++s) {
            int a = ora_sample_action(g->seed, g->env_id, step0 + s, ora_game_valid_mask(g, -1), weights);
            step_one(b, g, a);
            ++total;
        }
    }
    return total;
}

void ora_batch_export(ora_batch *b, double *credits, double *bets, double *pending, double *payoffs, uint8_t *states,
                      double *minraise, int32_t *turn, int32_t *hand, int32_t *dealer, int32_t *sb, int32_t *bb,
                      int32_t *active, uint32_t *episode)
{
    size_t n = b->n;
    for (size_t i = 0; i < n; ++i) {
        const ora_game *g = &b->games[i];
        for (int p = 0; p < g->num_players; ++p) {
            if (credits) credits[(size_t)p * n + i] = g->credits[p];
            if (bets) bets[(size_t)p * n + i] = g->bets[p];
            if (pending) pending[(size_t)p * n + i] = g->pending[p];
            if (payoffs) payoffs[(size_t)p * n + i] = g->payoffs[p];
            if (states) states[(size_t)p * n + i] = g->states[p];
        }
        if (minraise) minraise[i] = g->minimum_raise_value;
        if (turn) turn[i] = g->turn;
        if (hand) hand[i] = g->hand;
        if (dealer) dealer[i] = g->dealer;
        if (sb) sb[i] = g->small_blind_idx;
        if (bb) bb[i] = g->big_blind_idx;
        if (active) active[i] = g->active_player;
        if (episode) episode[i] = g->episode;
    }
}
