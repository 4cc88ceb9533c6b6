"""Numeric constants of the reference (pokerl/enums.py:1-138); values must match exactly
because they are baked into the CUDA kernels (pokerl_b200/csrc/pkb_common.cuh)."""


class HandRanking:
    """Hand categories, lower is better (reference enums.py:9-33)."""
    STRAIGHT_FLUSH, POKER, FULL, FLUSH, STRAIGHT, TRIS, TWO_PAIR, PAIR, HIGH, NONE = range(1, 11)

    as_string = ['Five of a Kind', 'Straight Flush', 'Four of a Kind', 'Full House', 'Flush', 'Straight',
                 'Three of a Kind', 'Two Pair', 'One Pair', 'High Card', 'Nothing']


class CardSuit:
    """Suits (reference enums.py:38-54)."""
    SPADES, HEARTS, DIAMONDS, CLUBS = range(4)
    NUM_SUITS = 4
    MAX_SUIT = 3

    as_string = ['Spades', 'Hearts', 'Diamonds', 'Clubs']
    as_symbol = ['♠', '♥', '♦' '♣']  # sic: the reference concatenates the last two
    from_symbol = {'S': SPADES, 'H': HEARTS, 'D': DIAMONDS, 'C': CLUBS}


class CardRank:
    """Ranks; ONE and ACE are the same card, ACE (13) is its value when ranking (reference enums.py:59-95)."""
    ONE, TWO, THREE, FOUR, FIVE, SIX, SEVEN, EIGHT, NINE, TEN, JACK, QUEEN, KING, ACE = range(14)
    NUM_RANKS = 13
    MAX_RANK = 13

    as_string = ['Ace', 'Two', 'Three', 'Four', 'Five', 'Six', 'Seven', 'Eight', 'Nine', 'Ten', 'Jack', 'Queen',
                 'King', 'Ace']
    as_symbol = ['1', '2', '3', '4', '5', '6', '7', '8', '9', 'T', 'J', 'Q', 'K', 'A']
    from_symbol = {sym: rank for rank, sym in enumerate(as_symbol)}


class PokerMoves:
    """Discrete action set (reference enums.py:104-116)."""
    FOLD, CHECK, CALL = 0, 1, 2
    RAISE_ANY = RAISE_TEN = 3
    RAISE_QUARTER, RAISE_HALF, ALL_IN = 4, 5, 6
    NUM_MOVES = 7
    NUM_RAISE_MOVES = 3

    as_string = ['Fold', 'Check', 'Call', 'Raise 10%', 'Raise 25%', 'Raise half', 'All-in']


class PlayerState:
    """Per-seat state (reference enums.py:130-138)."""
    FOLDED, ACTIVE, CALLED, ALL_IN, BROKEN = range(5)
    NUM_STATES = 5

    as_string = ['Folded', 'Active', 'Called', 'All-in', 'Broken']
