// hostcheck.cpp -- compiles the PRODUCT's host/device headers (pokerl_b200/csrc/*.cuh) for the host so
// the CPU test-suite can check their algorithms against the oracle without a GPU.  Test-only: the
// product never loads this library, and the GPU tests check the same functions through the C ABI.
#include <stddef.h>
#include <stdint.h>

#include "pkb_eval.cuh"
#include "pkb_philox.cuh"

// the closed form of the winner scan used by the step kernels' side-pot loop: winner mask of n (2..10) packed rankings
template <int P>
static uint32_t closed(const uint32_t* packed)
{
    uint32_t rk[P];
    for (int i = 0; i < P; ++i) rk[i] = packed[i];
    return pkb::winners_closed<P>(rk);
}

extern "C" {

uint32_t hc_eval_mask(uint64_t m) { return pkb::eval_mask52(m); }
// the table-driven form used by the batch kernels (table built with the product's own entry function)
static const uint32_t* eval_table()
{
    static uint32_t tbl[pkb::EVAL_TABLE_SIZE];
    static bool ready = false;
    if (!ready) {
        for (uint32_t x = 0; x < pkb::EVAL_TABLE_SIZE; ++x) tbl[x] = pkb::eval_table_entry(x);
        ready = true;
    }
    return tbl;
}
uint32_t hc_eval_mask_tbl(uint64_t m) { return pkb::eval_mask52_tbl(m, eval_table()); }
// number of k-card hands (all C(52,k), in increasing mask order) on which the two forms differ
uint64_t hc_eval_tbl_mismatches(int k, uint64_t count)
{
    const uint32_t* tbl = eval_table();
    uint64_t m = (1ull << k) - 1, bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        bad += pkb::eval_mask52_tbl(m, tbl) != pkb::eval_mask52(m);
        const uint64_t c = m & -m, r = m + c;
        m = (((r ^ m) >> 2) / c) | r;
    }
    return bad;
}
// the select form used inside the step kernels against the bitmask evaluator, same enumeration
uint32_t hc_eval_mask_sel(uint64_t m) { return pkb::eval_mask52_sel(m); }
uint64_t hc_eval_sel_mismatches(int k, uint64_t first_index_unused, uint64_t count)
{
    (void)first_index_unused;
    uint64_t m = (1ull << k) - 1, bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        bad += pkb::eval_mask52_sel(m) != pkb::eval_mask52(m);
        const uint64_t c = m & -m, r = m + c;
        m = (((r ^ m) >> 2) / c) | r;
    }
    return bad;
}
// the perfect-hash seven-card evaluator of the batch kernels, tables built with the product's own host code: -1 if the keys collide, -2 if the displacement hash cannot be built
long long hc_eval_hash3_mismatches(uint64_t count)
{
    static uint32_t* nf = nullptr;
    static uint32_t klo[pkb::EVAL_KLO_SIZE], khi[pkb::EVAL_KHI_SIZE], flt[8192];
    static uint16_t* disp = nullptr;
    static uint32_t* dense = nullptr;
    if (!nf) {
        nf = new uint32_t[pkb::EVAL_NF_SIZE];
        if (pkb::eval_nf_table_fill(nf) != 49205u) return -1;
        disp = new uint16_t[pkb::EVAL_DISP_SIZE];
        dense = new uint32_t[pkb::EVAL_DENSE_SIZE];
        if (!pkb::eval_dense_table_fill(nf, disp, dense)) return -2;
        for (uint32_t x = 0; x < (uint32_t)pkb::EVAL_KLO_SIZE; ++x) klo[x] = pkb::raw_key_sum(x);
        for (uint32_t x = 0; x < (uint32_t)pkb::EVAL_KHI_SIZE; ++x) khi[x] = pkb::raw_key_sum(x << 7);
        for (uint32_t x = 0; x < 8192u; ++x) flt[x] = pkb::flush_table_entry(pkb::rot_ranks(x));
    }
    uint64_t m = (1ull << 7) - 1;
    long long bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        bad += pkb::eval7_hash3(m, klo, khi, disp, dense, flt) != pkb::eval_mask52(m);
        const uint64_t c = m & -m, r = m + c;
        m = (((r ^ m) >> 2) / c) | r;
    }
    return bad;
}
// the streaming form (eval7_fast): every seven-card hand is either marked rare or its table value equals the bitmask
// evaluator; returns mismatches, *rare_count = hands marked rare; -1 / -2 as above.  Masks with other card counts
// (k != 7, first `other` masks of each k in 0..8) must all be marked rare: -3 if one is not.
long long hc_eval_fast_mismatches(uint64_t count, uint64_t* rare_count, uint64_t other)
{
    static uint32_t* nf = nullptr;
    static uint32_t klo[128], khi[64];
    static uint8_t* disp = nullptr;
    static uint32_t* dense = nullptr;
    if (!nf) {
        nf = new uint32_t[pkb::EVAL_NF_SIZE];
        if (pkb::eval_nf_table_fill(nf, true) != 49205u) return -1;
        disp = new uint8_t[pkb::EVAL_FAST_BUCKETS];
        dense = new uint32_t[pkb::EVAL_DENSE_SIZE];
        if (!pkb::eval_fast_table_fill(nf, disp, dense)) return -2;
        for (uint32_t x = 0; x < 128u; ++x) klo[x] = pkb::fast_klo_entry(x);
        for (uint32_t x = 0; x < 64u; ++x) khi[x] = pkb::fast_khi_entry(x);
    }
    uint64_t m = (1ull << 7) - 1, rares = 0;
    long long bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        uint32_t rare = 0;
        const uint32_t v = pkb::eval7_fast(m, klo, khi, 1, 0, disp, dense, &rare);
        if (rare) ++rares;
        else bad += v != pkb::eval_mask52(m);
        const uint64_t c = m & -m, r = m + c;
        m = (((r ^ m) >> 2) / c) | r;
    }
    for (int k = 0; k <= 8; ++k) {
        if (k == 7) continue;
        uint64_t mk = (1ull << k) - 1;
        for (uint64_t i = 0; i < other; ++i) {
            uint32_t rare = 0;
            (void)pkb::eval7_fast(mk, klo, khi, 1, 0, disp, dense, &rare);
            if (!rare) return -3;
            if (k == 0) break;
            const uint64_t c = mk & -mk, r = mk + c;
            mk = ((((r ^ mk) >> 2) / c) | r) & ((1ull << 52) - 1);
            if (!mk || pkb::popc64(mk) != k) break;
        }
    }
    if (rare_count) *rare_count = rares;
    return bad;
}
uint32_t hc_eval_cards(const uint8_t* c, int n) { return pkb::eval_cards(c, n); }
uint32_t hc_eval_cards_literal(const uint8_t* c, int n) { return pkb::eval_cards_literal(c, n); }

void hc_eval_masks(const uint64_t* m, size_t n, uint32_t* out)
{
    for (size_t i = 0; i < n; ++i) out[i] = pkb::eval_mask52(m[i]);
}

static uint64_t next_same_popcount(uint64_t v)
{
    uint64_t c = v & -v, r = v + c;
    return (((r ^ v) >> 2) / c) | r;
}

// `count` masks of k cards in increasing order starting at `start` (0: the lowest, (1 << k) - 1); returns the next one
uint64_t hc_enumerate_masks(int k, uint64_t start, uint64_t count, uint64_t* out)
{
    uint64_t m = start ? start : (1ull << k) - 1;
    for (uint64_t i = 0; i < count; ++i) {
        out[i] = m;
        m = next_same_popcount(m);
    }
    return m;
}

// all C(52,k) masks in increasing order starting from the lowest; hist/sums as pkb_eval_exhaustive
void hc_eval_exhaustive(int k, uint64_t* hist, uint64_t* sums, uint32_t* out, uint64_t count)
{
    uint64_t m = (1ull << k) - 1;
    for (uint64_t i = 0; i < count; ++i) {
        const uint32_t p = pkb::eval_mask52(m);
        hist[p >> 20]++;
        sums[0] += p;
        sums[1] += (uint64_t)p * (m % 1000003ull);
        if (out) out[i] = p;
        m = next_same_popcount(m);
    }
}

void hc_deal(uint64_t seed, uint32_t env, uint32_t episode, uint32_t hand, int P, uint8_t* out)
{
    pkb::Dealer d(seed, env, episode, hand);
    for (int i = 0; i < 5 + 2 * P; ++i) out[i] = (uint8_t)pkb::card_value_from_id(d.next());
}

// the all-at-once deal of the step kernels (small tables): must equal the sequential Dealer
int hc_deal_ids(uint64_t seed, uint32_t env, uint32_t episode, uint32_t hand, int P, uint8_t* out)
{
    uint32_t id[13];
    switch (P) {
        case 1: { uint32_t a[7]; pkb::deal_ids<7>(seed, env, episode, hand, a); for (int i = 0; i < 7; ++i) id[i] = a[i]; } break;
        case 2: { uint32_t a[9]; pkb::deal_ids<9>(seed, env, episode, hand, a); for (int i = 0; i < 9; ++i) id[i] = a[i]; } break;
        case 3: { uint32_t a[11]; pkb::deal_ids<11>(seed, env, episode, hand, a); for (int i = 0; i < 11; ++i) id[i] = a[i]; } break;
        case 4: { uint32_t a[13]; pkb::deal_ids<13>(seed, env, episode, hand, a); for (int i = 0; i < 13; ++i) id[i] = a[i]; } break;
        default: return -1;
    }
    for (int i = 0; i < 5 + 2 * P; ++i) out[i] = (uint8_t)pkb::card_value_from_id(id[i]);
    return 0;
}

int hc_sample(uint64_t seed, uint32_t env, uint32_t step, uint32_t valid, const uint32_t* w)
{
    return pkb::sample_action(seed, env, step, valid, w);
}

uint32_t hc_winners_closed(const uint32_t* packed, int n)
{
    switch (n) {
        case 2: return closed<2>(packed);
        case 3: return closed<3>(packed);
        case 4: return closed<4>(packed);
        case 5: return closed<5>(packed);
        case 6: return closed<6>(packed);
        case 7: return closed<7>(packed);
        case 8: return closed<8>(packed);
        case 9: return closed<9>(packed);
        case 10: return closed<10>(packed);
    }
    return 0xffffffffu;
}
uint32_t hc_winners_scan(const uint32_t* packed, int n)
{
    pkb::WinnerScan ws;
    for (int i = 0; i < n; ++i) ws.feed(i, packed[i]);
    return ws.winners;
}
// position of the j-th set bit: the branch-free form against the out-of-line one
uint32_t hc_select_bit_mismatches(uint64_t seed, uint64_t count)
{
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1, bad = 0;
    for (uint64_t i = 0; i < count; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const uint64_t f = x & ((1ull << 52) - 1);
        const int n = pkb::popc64(f);
        for (int j = 0; j < n; ++j) bad += pkb::select_bit_inl(f, (uint32_t)j) != pkb::select_bit(f, (uint32_t)j);
    }
    return (uint32_t)bad;
}

int hc_compare(const uint32_t* packed, int n, uint8_t* onehot)
{
    pkb::WinnerScan ws;
    for (int i = 0; i < n; ++i) ws.feed(i, packed[i]);
    for (int i = 0; i < n; ++i) onehot[i] = (ws.winners >> i) & 1u;
    return pkb::popc32(ws.winners);
}
}
