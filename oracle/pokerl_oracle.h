/*
 * pokerl_oracle.h -- CPU oracle for the pokerl hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a scalar, plain-C restatement of the reference algorithm
 * (sneppy/pokerl v0.0.6: pokerl/judger.py, pokerl/game.py, pokerl/cards.py).
 * It exists to CHECK the CUDA product path and to serve as bench.py's
 * cpu_baseline leg.  Nothing under pokerl_b200/ may include, link or call it.
 *
 * Parity pins (see oracle/README.md): the 22+7 known-answer vectors of the
 * reference's tests/pokerl/test_judger.py, the exhaustive K=5/6/7 golden sums
 * of BASELINE.md section 3 (computed from the reference's judger.py), and
 * recorded step traces of the live reference Game (tests/golden/).
 *
 * Card encoding (cards.py:14-26): value = (suit << 4) | rank0, rank0 0=Ace,
 * 1=Two .. 12=King; rank = rank0 or 13; id = rank0 + 13*suit.
 */
#ifndef POKERL_ORACLE_H
#define POKERL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAX_PLAYERS 23 /* 5 + 2P <= 52 */
#define ORA_NUM_MOVES 7

/* enums.py:130-134 */
enum { ORA_FOLDED = 0, ORA_ACTIVE = 1, ORA_CALLED = 2, ORA_ALL_IN = 3, ORA_BROKEN = 4 };
/* enums.py:104-111 */
enum { ORA_FOLD = 0, ORA_CHECK = 1, ORA_CALL = 2, ORA_RAISE_TEN = 3, ORA_RAISE_QUARTER = 4, ORA_RAISE_HALF = 5, ORA_MOVE_ALL_IN = 6 };
/* step flag bits (this repo's ABI, mirrored by the CUDA path) */
enum { ORA_F_GAME_OVER = 1, ORA_F_HAND_OVER = 2, ORA_F_ROUND_OVER = 4, ORA_F_SKIPPED = 32, ORA_F_INVALID = 64, ORA_F_POISONED = 128 };
#define ORA_ACTION_SKIP (-2147483647 - 1) /* ora_batch_step leaves that env alone (masked lockstep advance) */

#define ORA_NEXT_TURN_CAP 64

/* ---- judger.py ---- */
/* packed = (category << 20) | get_kickers_value(kickers)   (judger.py:7-109) */
uint32_t ora_eval_cards(const uint8_t *cards, int n);
void ora_eval_cards_batch(const uint8_t *cards, int k, size_t n, uint32_t *out);
/* mask bit i = card with Card.id == i (cards.py:22-26) */
uint32_t ora_eval_mask(uint64_t mask);
void ora_eval_masks_batch(const uint64_t *masks, size_t n, uint32_t *out);
/* all C(52,k) subsets; hist[cat], sumA = sum packed, sumB = sum packed*(mask % 1000003) (mod 2^64) */
void ora_eval_exhaustive(int k, uint64_t hist[11], uint64_t *sumA, uint64_t *sumB);
/* writes packed values of all C(52,k) hands in colex order of the id-mask (same order as the CUDA enumerator) */
void ora_eval_enumerate(int k, uint64_t first, uint64_t count, uint32_t *out);
/* judger.py:111-158; returns the number of winners, onehot[n] filled */
int ora_compare_rankings(const uint32_t *packed, int n, uint8_t *onehot);

/* ---- dealing and sampling (this repo's Philox spec, not the reference's MT19937) ---- */
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* cards[0..4] board, cards[5+2p..6+2p] hole of seat p; values are (suit<<4)|rank0 */
void ora_deal(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t hand, int num_players, uint8_t *cards);
/* weighted choice over the valid mask (bit a = action a valid); weights NULL = uniform */
int ora_sample_action(uint64_t seed, uint32_t env_id, uint32_t step_idx, unsigned valid_mask, const uint32_t *weights);

/* numpy float64 add.reduce order (n<8 sequential, n>=8 eight-lane pairwise) */
double ora_np_sum(const double *a, int n);

/* ---- game.py ---- */
typedef struct ora_game {
    int num_players;
    double start_credits[ORA_MAX_PLAYERS];
    double big_blind, small_blind;
    int cfg_dealer;
    uint8_t states[ORA_MAX_PLAYERS];
    double credits[ORA_MAX_PLAYERS], bets[ORA_MAX_PLAYERS], pending[ORA_MAX_PLAYERS], payoffs[ORA_MAX_PLAYERS];
    double minimum_raise_value;
    int turn, hand, dealer, small_blind_idx, big_blind_idx, active_player;
    uint8_t deck[5 + 2 * ORA_MAX_PLAYERS];
    /* dealing */
    uint64_t seed;
    uint32_t env_id, episode;
    const uint8_t *deck_queue; /* optional injected decks, deck_stride bytes each */
    int deck_stride, deck_count, deck_cursor;
    /* per-step diagnostics */
    int next_turn_calls, end_hand_calls, showdown_layers, poisoned;
} ora_game;

void ora_game_init(ora_game *g, int num_players, const double *start_credits, double big_blind, double small_blind,
                   int dealer, uint64_t seed, uint32_t env_id);
void ora_game_set_decks(ora_game *g, const uint8_t *decks, int stride, int count);
void ora_game_reset(ora_game *g);
unsigned ora_game_valid_mask(const ora_game *g, int player); /* player < 0: active player */
int ora_game_is_over(const ora_game *g);
/* returns flag bits; invalid action leaves the state untouched */
unsigned ora_game_step(ora_game *g, int action);

/* ---- batch of N lockstep games (SoA export mirrors the CUDA layout) ---- */
typedef struct ora_batch ora_batch;
ora_batch *ora_batch_create(size_t n, int num_players, const double *start_credits, double big_blind, double small_blind,
                            int dealer, uint64_t seed, uint32_t env_id_base, int auto_reset);
void ora_batch_destroy(ora_batch *b);
void ora_batch_set_threads(int nthreads);
void ora_batch_reset(ora_batch *b);
void ora_batch_step(ora_batch *b, const int32_t *actions, uint8_t *flags, uint8_t *valid);
void ora_batch_valid(ora_batch *b, uint8_t *valid);
void ora_batch_sample(ora_batch *b, uint32_t step_idx, const uint32_t *weights, int32_t *actions);
/* fused sample+step for T steps (the cpu_baseline loop); returns total env-steps executed */
uint64_t ora_batch_run_random(ora_batch *b, uint32_t step0, uint32_t steps, const uint32_t *weights);
/* any pointer may be NULL; f64 arrays are [P][N], states u8 [P][N], the rest [N] */
void ora_batch_export(ora_batch *b, double *credits, double *bets, double *pending, double *payoffs, uint8_t *states,
                      double *minraise, int32_t *turn, int32_t *hand, int32_t *dealer, int32_t *sb, int32_t *bb,
                      int32_t *active, uint32_t *episode);
ora_game *ora_batch_game(ora_batch *b, size_t i);

#ifdef __cplusplus
}
#endif
#endif
