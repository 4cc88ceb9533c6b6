"""examples/random_game.py of the reference (lines 1-13), unchanged except for the import: one 4-player
game with random agents through the scalar `Game` API.  Needs a B200 (there is no CPU fallback)."""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pokerl_b200 import Game  # noqa: E402
from pokerl_b200.agents import RandomAgent  # noqa: E402

if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    agent = RandomAgent()
    game = Game(num_players=4, start_credits=1000, big_blind=40, small_blind=20)  # random_game.py:9
    game.reset()
    done, steps = False, 0
    while not done:
        done, hand_over, _ = game.step(agent(game.active_state))
        steps += 1
        if hand_over:
            logging.info("hand %d over after %d steps; credits %s", game.hand - 1, steps, game.credits)
    print("game over after", steps, "steps; credits", game.credits)
