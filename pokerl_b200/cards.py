"""Card and deck helpers with the reference's encodings (pokerl/cards.py:4-77)."""
from typing import List, Tuple, Union

from .enums import CardRank, CardSuit


class Card:
    """A playing card; `value = (suit << 4) | rank0` with rank0 0 = ace .. 12 = king (cards.py:4-26)."""

    __slots__ = ("value",)

    def __init__(self, value: Union[int, str, Tuple[int, int]]):
        if isinstance(value, tuple):  # (rank, suit); rank 13 and rank 0 are both the ace (cards.py:47-51)
            rank, suit = value
            value = (suit << 4) | (CardRank.ONE if rank == CardRank.ACE else rank)
        elif isinstance(value, str):  # 'RS', e.g. 'TH', '1D', 'AS' (cards.py:52-56)
            rank, suit = CardRank.from_symbol[value[0]], CardSuit.from_symbol[value[1]]
            value = (suit << 4) | (CardRank.ONE if rank == CardRank.ACE else rank)
        assert (value & 0xf) < CardRank.NUM_RANKS, 'Invalid card rank'
        assert (value >> 4) < CardSuit.NUM_SUITS, 'Invalid card suit'
        self.value = int(value)

    @property
    def rank(self) -> int:
        """Rank with the ace high: 1 = two .. 13 = ace (cards.py:8-14)."""
        return (self.value & 0xf) or 13

    @property
    def suit(self) -> int:
        return self.value >> 4

    @property
    def id(self) -> int:
        """Unique id in [0, 52): rank0 + 13 * suit (cards.py:22-26); bit index of hand masks."""
        return (self.value & 0xf) + (self.value >> 4) * 13

    @classmethod
    def from_id(cls, card_id: int) -> "Card":
        return cls(((card_id // 13) << 4) | (card_id % 13))

    # pickled like the reference's plain-attribute class (state = {'value': v}), so either side can load the other's cards
    def __getstate__(self):
        return {"value": self.value}

    def __setstate__(self, state):
        if isinstance(state, tuple):  # (dict or None, slot dict)
            state = {**(state[0] or {}), **(state[1] or {})}
        self.value = int(state["value"])

    def __eq__(self, other):
        return isinstance(other, Card) and other.value == self.value

    def __hash__(self):
        return hash(self.value)

    def __repr__(self) -> str:
        # Unicode playing cards; the block has a Knight between Jack and Queen (cards.py:66-72)
        return chr((0x1f0a2 if (self.value & 0xf) > CardRank.JACK else 0x1f0a1) + self.value)


def create_default_deck() -> List[Card]:
    """The 52 cards in the reference's order: rank-major, suits inside (cards.py:74-77)."""
    return [Card((rank, suit)) for rank in range(CardRank.NUM_RANKS) for suit in range(CardSuit.NUM_SUITS)]
