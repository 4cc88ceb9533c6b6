"""Known-answer vectors of the reference's own test-suite (tests/pokerl/test_judger.py), as data.

EVAL_KATS:    (cards, (category, kickers))        test_judger.py:14-78   (22 vectors)
COMPARE_KATS: ([hand, ...], onehot winners)       test_judger.py:83-117  (7 vectors)
Card strings are rank symbol + suit symbol; '1' and 'A' are both the ace (enums.py:80-95).
Four eval vectors and one compare vector contain duplicate cards, exactly as upstream.
"""
SF, POKER, FULL, FLUSH, STRAIGHT, TRIS, TWO_PAIR, PAIR, HIGH = range(1, 10)
TWO, THREE, FOUR, FIVE, SIX, SEVEN, EIGHT, NINE, TEN, JACK, QUEEN, KING, ACE = range(1, 14)

EVAL_KATS = [
    ("2C 3C 4C 5C 6C 7C 8C", (SF, [EIGHT])),                               # :14
    ("1D 2D 3D 4D 5D 6C 7C", (SF, [FIVE])),                                # :17
    ("1D 2D 3D 4D 8D 6C 7C", (FLUSH, [ACE, EIGHT, FOUR, THREE, TWO])),     # :20
    ("1D 3D 5H 7H 7S 9S JS", (PAIR, [SEVEN, ACE, JACK, NINE])),            # :23
    ("1D 3D 5H 8H 7S 9S 1S", (PAIR, [ACE, NINE, EIGHT, SEVEN])),           # :26
    ("1D 1C 5H 6H 6H TS TS", (TWO_PAIR, [ACE, TEN, SIX])),                 # :29 (duplicates)
    ("KD QC JH TH 9H 9S 9D", (STRAIGHT, [KING])),                          # :32
    ("KD QC JH TH 9H 5S 9D", (STRAIGHT, [KING])),                          # :35
    ("KD QC JH TH 9H 5S 7D", (STRAIGHT, [KING])),                          # :38
    ("8D 9C 7H 6H 9H 5S 7D", (STRAIGHT, [NINE])),                          # :41
    ("2H 3C 4H 5H 6H 7H 8D", (FLUSH, [SEVEN, SIX, FIVE, FOUR, TWO])),      # :44
    ("8D 8C 8H 5H 5S 5D 3D", (FULL, [EIGHT, FIVE])),                       # :47
    ("8D 8C 8H 5H 5S 4D 3D", (FULL, [EIGHT, FIVE])),                       # :50
    ("1D 1C 1H 5H 5S 5D 5C", (POKER, [FIVE, ACE])),                        # :53
    ("1D 1C 1H 1S KS KD KC", (POKER, [ACE, KING])),                        # :56
    ("2D 3D 4D 5D 7D 6C KC", (FLUSH, [SEVEN, FIVE, FOUR, THREE, TWO])),    # :59
    ("2D 3C 4D 5D 8D JC KC", (HIGH, [KING, JACK, EIGHT, FIVE, FOUR])),     # :62
    ("2D 2C 3D 3D JD JC KC", (TWO_PAIR, [JACK, THREE, KING])),             # :65 (duplicates)
    ("2D 2C 3D 3D JD JC JC", (FULL, [JACK, THREE])),                       # :68 (duplicates)
    ("2D 2C 3D 3D 3D JC JC", (FULL, [THREE, JACK])),                       # :71 (duplicates)
    ("KC QC JC TC 9C AD 4D", (SF, [KING])),                                # :74
    ("9D 8C 4D 5D 6D JD KC", (FLUSH, [JACK, NINE, SIX, FIVE, FOUR])),      # :77
]

COMPARE_KATS = [
    (["1D 1C 1H 5H 6H 1S KS", "1D 1C 1H 5H 6H QS JS"], [1, 0]),            # :83
    (["1D 1C 1H 5H 6H 1S KS", "1D 1C 1H 5H 6H QH JH"], [1, 0]),            # :88
    (["1D 1C 1H 2H 3H 1S KS", "1D 1C 1H 2H 3H 4H 5H"], [0, 1]),            # :93
    (["1D 3C 5H 7H 9H KS JS", "1D 3C 5H 7H 9H QS TS"], [1, 0]),            # :98
    (["1D 3C 5H 7H 9H KS JS", "1D 3C 5H 7H 9H QS TS"], [1, 0]),            # :103
    (["1D 3C 5H 7H 9H KS JS", "1D 3C 5H 7H 9H 3S TS"], [0, 1]),            # :108
    (["1D 3C 5H 7H 9H 1S 9D", "1D 3C 5H 7H 9H 7S 7S", "1D 3C 5H 7H 9H 9S 9S"], [0, 0, 1]),  # :113 (duplicates)
]

_RANK = {s: i for i, s in enumerate("123456789TJQK")}
_RANK["A"] = 0
_SUIT = {"S": 0, "H": 1, "D": 2, "C": 3}


def card_value(sym: str) -> int:
    """'TH' -> (suit << 4) | rank0, the reference's Card('RS') constructor (cards.py:52-56)."""
    return (_SUIT[sym[1]] << 4) | _RANK[sym[0]]


def hand_values(s: str):
    return [card_value(x) for x in s.split()]


# golden sums computed from the reference's judger.py over all C(52,k) hands
EXHAUSTIVE_GOLDEN = {
    5: ([0, 40, 624, 3744, 5108, 10194, 54912, 123552, 1098240, 1302546, 0], 23934246850430, 11940267641381718592),
    6: ([0, 1732, 14664, 165984, 205568, 361798, 732160, 2532816, 9730800, 6612998, 0], 176534665307528,
        14502619058268878659),
    7: ([0, 36768, 224848, 3473184, 4032468, 6197789, 6461665, 31433660, 58628962, 23295216, 0], 1068968016237707,
        18108495491170714949),
}

# Random123 known-answer vectors for Philox4x32-10: (counter, key, output)
PHILOX_KATS = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]
