"""`Game.StateView` on the wire (reference game.py:208-240; pokerl/network/server.py sends it pickled).

The reference pickles a StateView as the 10-tuple of `__getstate__`; this package keeps the same tuple.  What differs
is only the module path recorded in the pickle, so:

* `loads(data)` reads a pickle written by EITHER package (`pokerl.game.Game.StateView`, `pokerl.cards.Card` are mapped
  to the classes of this package);
* `dumps(view)` writes a scalar-mode StateView so that the unmodified reference can `pickle.loads` it: protocol 2,
  with the globals renamed to the reference's modules.

The socket layer itself (pokerl/network) is out of scope.
"""
import io
import pickle

_TO_OURS = {("pokerl.game", "Game.StateView"): ("pokerl_b200.game", "Game.StateView"),
            ("pokerl.game", "Game"): ("pokerl_b200.game", "Game"),
            ("pokerl.cards", "Card"): ("pokerl_b200.cards", "Card")}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        module, name = _TO_OURS.get((module, name), (module, name))
        return super().find_class(module, name)


def loads(data: bytes):
    """Load a pickled StateView (or any object built from StateViews / Cards) written by the reference or by this package."""
    return _Unpickler(io.BytesIO(data)).load()


def dumps(view) -> bytes:
    """Pickle a scalar-mode StateView so that `pickle.loads` in a process that has the reference installed yields a
    `pokerl.game.Game.StateView`."""
    if getattr(view, "_batched", False):
        raise ValueError("only scalar-mode StateViews have a reference wire format (tensors do not travel as numpy arrays)")
    data = pickle.dumps(view, protocol=2)  # GLOBAL opcodes are newline-delimited text: safe to rename
    return data.replace(b"cpokerl_b200.game\n", b"cpokerl.game\n").replace(b"cpokerl_b200.cards\n", b"cpokerl.cards\n")
