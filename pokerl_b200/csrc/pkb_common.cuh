// pkb_common.cuh -- shared definitions for the pokerl_b200 CUDA path (sm_100a).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PKB_HD __host__ __device__ __forceinline__
#define PKB_HDM __host__ __device__ __forceinline__ /* member functions */
#define PKB_D __device__ __forceinline__
#define PKB_HD_CALL __host__ __device__ __noinline__ /* real calls: keeps the step kernel's code instruction-cache sized */
#else
#define PKB_HD static inline
#define PKB_HDM inline
#define PKB_D static inline
#define PKB_HD_CALL static inline
#endif

namespace pkb {

// enums.py:130-134
enum : int { ST_FOLDED = 0, ST_ACTIVE = 1, ST_CALLED = 2, ST_ALL_IN = 3, ST_BROKEN = 4 };
// enums.py:104-111
enum : int { MV_FOLD = 0, MV_CHECK = 1, MV_CALL = 2, MV_RAISE = 3, MV_ALL_IN = 6, NUM_MOVES = 7 };
// enums.py:9-18
enum : uint32_t { HR_STRAIGHT_FLUSH = 1, HR_POKER = 2, HR_FULL = 3, HR_FLUSH = 4, HR_STRAIGHT = 5, HR_TRIS = 6,
                  HR_TWO_PAIR = 7, HR_PAIR = 8, HR_HIGH = 9, HR_NONE = 10 };
constexpr uint32_t PACKED_NONE = HR_NONE << 20;

PKB_HD int popc32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
PKB_HD int popc64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// index of the highest set bit; x != 0
PKB_HD int topbit32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
// 1-based index of the lowest set bit, 0 for x == 0
PKB_HD int ffs32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}
PKB_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// card id (rank0 + 13*suit, cards.py:22-26) -> card value ((suit << 4) | rank0, cards.py:4-20)
PKB_HD uint32_t card_value_from_id(uint32_t id)
{
    uint32_t suit = (id * 20165u) >> 18;  // id / 13 for id < 64
    return (suit << 4) | (id - 13u * suit);
}
PKB_HD uint32_t card_id_from_value(uint32_t v) { return (v & 0xfu) + 13u * (v >> 4); }

}  // namespace pkb
