"""Headless mirrors of the reference's three driver loops -- the callers of the hot path.

The originals cannot be imported on a GPU box (pygame / tensorflow / imageio at module
top: self_play.py:4, choose_best_player.py:4-6, main.py:5), so the loops are restated here
against this package's ``Player`` / ``ResNet`` / ``utils``; each function names the lines it
follows.  ``Arena`` is the lock-step form of ``choose_best_player.py:42-72`` (BASELINE
config 5): all games of a match run at once on the device, one private table per player
per game, and the reference's sequential win counting -- draws skipped, early stop after
30 games -- is applied to the results in game order.
"""
from __future__ import annotations

import numpy as np
import torch

from . import rules, utils
from .engine import SearchEngine, make_config
from .genData.player import Player, board_to_state


# ------------------------------------------------------------------------------------
# self_play.py:94-101 -- one AI-vs-AI game, training=False, deterministic best move
# ------------------------------------------------------------------------------------
def self_play_game(config, pv_fn, max_plies=None):
    """Returns (moves [(i, j)...], final state string, value) of one game played by
    ``Player(config, training=False, pv_fn=pv_fn)``.  As in the viewer, ``last_action`` is
    reset to ``None`` before every call (self_play.py:95-97), so the third input plane is
    always empty."""
    player = Player(config, training=False, pv_fn=pv_fn)
    state = player.get_init_state()
    game_over, value, moves = False, 0.0, []
    while not game_over and (max_plies is None or len(moves) < max_plies):
        action = None
        _, action = player.get_action(state, last_action=action)
        board = utils.step(utils.state_to_board(state, config.board_size), action)
        state = utils.board_to_state(board)
        game_over, value = utils.is_game_over(board, config.goal)
        moves.append(action)
    player.close()
    return moves, state, value


# ------------------------------------------------------------------------------------
# main.py:82-94 -- the data-generating worker
# ------------------------------------------------------------------------------------
def label_result(game_record) -> int:
    """main.py:86-93: DRAW if the last value is 0, else BLACK_WIN for odd length."""
    value = game_record[-1][-2]
    if value == 0.0:
        return utils.DRAW
    return utils.BLACK_WIN if len(game_record) % 2 == 1 else utils.WHITE_WIN


def gen_data(pipe, q, config=None, games=None, seed=0):
    """One reference-style worker: ``Player(config, training=True, pipe=pipe).run()`` forever
    (or ``games`` times), each game put on ``q`` as ``(record, result)``.

    The worker creates a CUDA engine, and CUDA cannot be used in a *forked* child of a process that has
    already initialised it (main.py:50-55 forks after building the network): start these workers with
    ``multiprocessing.get_context("spawn").Process(target=gen_data, ...)``, or use ``gen_data_lockstep``
    (one process, thousands of games), which is what ``train_loop`` does."""
    import multiprocessing as mp
    if mp.current_process().name != "MainProcess" and mp.get_start_method(allow_none=True) in (None, "fork") \
            and torch.cuda.is_initialized():
        raise RuntimeError("gen_data: forked after CUDA was initialised in the parent; start the worker with the "
                           "'spawn' start method (or use gen_data_lockstep)")
    if config is None:
        from . import config as config_module
        config = config_module
    player = Player(config, training=True, pipe=pipe, seed=seed)
    k = 0
    while games is None or k < games:
        record = player.run()
        q.put((record, label_result(record)), block=True)
        k += 1
    player.close()


def gen_data_lockstep(selfplay, passes_per_poll=None):
    """Generator of ``(record, result)`` from a ``SelfPlay`` engine: the lock-step
    replacement of ``max_processes`` gen_data workers feeding ``q``."""
    k = passes_per_poll or selfplay.config.sims
    while True:
        selfplay.run_passes(k)
        for rec, result in selfplay.harvest_games():
            yield rec, result


# ------------------------------------------------------------------------------------
# main.py:29-78 -- the training loop, everything on the device
# ------------------------------------------------------------------------------------
def train_loop(config, n_games=4096, total_step=None, restore=None, seed=0, save_every=60, save_dir=None,
               passes_per_poll=None, log=print):
    """``main.main`` without the worker processes: lock-step self-play feeds the device-resident
    ``RandomStack``; every accepted game, once the buffer is full, is followed by four minibatches of
    ``config.batch_size`` (main.py:61-68) with the learning rate of ``config.get_lr(step)``; the new weights
    go back into the search net after every harvest (the reference's workers see every update through
    the shared session).  Returns (trainer, stack, step)."""
    import numpy as np
    from .genData.network import ResNet
    from .replay import parse_headers
    from .selfplay import SelfPlay
    from .train import Trainer
    from .utils import RandomStack
    S = config.board_size
    total_step = config.total_step if total_step is None else total_step
    stack = RandomStack(board_size=S, length=config.buffer_size)
    net = ResNet(S, max_batch=n_games, seed=seed)
    if restore:
        net.restore(restore)
    trainer = Trainer(S, {k: v.cpu().numpy() for k, v in net.device_net.params.items()})
    sp = SelfPlay(config, n_games=n_games, net=net.device_net, training=True, seed=seed)
    sp.start()
    step = 1
    while step < total_step:
        sp.run_passes(passes_per_poll or config.simulation_per_step)
        records, _ = sp.harvest()
        if records.shape[0] == 0:
            continue
        head = parse_headers(records)
        lens, res = head["game_len"], head["result"]
        i = 0
        while i < records.shape[0] and step < total_step:
            n = int(lens[i])
            accepted = stack.push(records[i:i + n], int(res[i]))      # main.py:61
            i += n
            if accepted and stack.is_full():                           # main.py:62
                for _ in range(4):
                    out = trainer.step(*stack.get_data_device(config.batch_size), lr=config.get_lr(step))
                step += 1
                log("step: %d, xcross_loss: %0.3f, mse: %0.3f, entropy: %0.3f" % (step, *out))
                if save_dir and step % save_every == 0:
                    trainer.sync_to(net)
                    net.save(f"{save_dir}/alphaFive", global_step=step)       # main.py:74
                    stack.save(step, directory=save_dir)
        trainer.sync_to(net)
    return trainer, stack, step


# ------------------------------------------------------------------------------------
# choose_best_player.py:42-72 -- arena between two weight sets
# ------------------------------------------------------------------------------------
def count_wins(winners, early_stop_after=30):
    """The reference's bookkeeping over games in index order: ``winners[i]`` is 0 / 1 for the
    winning player, -1 for a draw (skipped, :59-61).  From game 30 on the match stops as soon
    as one side has no win or the ratio leaves [0.5, 2] (:65-72).  Returns
    (wins0, wins1, games_counted)."""
    w = [0, 0]
    for i, who in enumerate(winners):
        if who < 0:
            continue
        w[who] += 1
        if i >= early_stop_after:
            if w[0] == 0 or w[1] == 0 or w[0] / w[1] > 2.0 or w[0] / w[1] < 0.5:
                return w[0], w[1], i + 1
    return w[0], w[1], len(winners)


class Arena:
    """``n_games`` simultaneous games between ``net0`` and ``net1`` (DeviceNet).  Game i is
    opened by player ``i % 2`` (choose_best_player.py:48); both players are
    ``Player(training=False)`` asked with ``random_a=True`` (:52), each searching only on its
    own turns in its own table."""

    def __init__(self, config, net0, net1, n_games, seed=0, game_id_base=0, check_every=16):
        self.config, self.N = config, n_games
        self.S = config.board_size
        self.nets = (net0, net1)
        mk = lambda s: SearchEngine(make_config(config, n_games=n_games, training=False, random_a=True,
                                                seed=s, game_id_base=game_id_base))
        self.engines = (mk(seed), mk(seed + 0x9E3779B9))
        self.check_every = check_every

    def play(self, max_plies=None):
        """Returns dict(winners int8[N] (0/1, -1 draw), plies int32[N], moves, leaf_evals)."""
        N, S = self.N, self.S
        dev = self.engines[0].device
        boards = torch.zeros((N, S, S), dtype=torch.int8, device=dev)
        last = torch.full((N,), -1, dtype=torch.int32, device=dev)
        mover = (torch.arange(N, device=dev) % 2).to(torch.int8)        # whose turn it is
        alive = torch.ones(N, dtype=torch.bool, device=dev)
        winners = torch.full((N,), -1, dtype=torch.int8, device=dev)
        plies = torch.zeros(N, dtype=torch.int32, device=dev)
        clear = torch.ones(N, dtype=torch.uint8, device=dev)             # both players reset() (:43-44)
        first = [True, True]
        ply = 0
        while bool(alive.any()) and (max_plies is None or ply < max_plies):
            action = torch.full((N,), -1, dtype=torch.int32, device=dev)
            for k in (0, 1):
                act_k = alive & (mover == k)
                if not bool(act_k.any()):
                    continue
                eng = self.engines[k]
                eng.set_roots(boards, last, act_k.to(torch.uint8), clear if first[k] else None)
                first[k] = False
                eng.run_search(net=self.nets[k], check_every=self.check_every, active=act_k)
                _, a = eng.finish_move()
                action = torch.where(act_k, a, action)
            nxt = rules.step(boards, action.clamp(min=0))
            codes = rules.terminal(nxt, self.config.goal)
            done = alive & (codes != 0)
            # the side to move at a terminal position has lost (code 2) or it is a draw (code 3)
            winners = torch.where(done & (codes == 2), mover, winners)
            plies = plies + alive.to(torch.int32)
            boards = torch.where(alive[:, None, None], nxt, boards)
            last = torch.where(alive, action, last)
            mover = torch.where(alive, 1 - mover, mover)
            alive = alive & ~done
            ply += 1
        c0, c1 = self.engines[0].counters(), self.engines[1].counters()
        return dict(winners=winners.cpu().numpy(), plies=plies.cpu().numpy(),
                    moves=c0["moves"] + c1["moves"], leaf_evals=c0["leaf_evals"] + c1["leaf_evals"],
                    final_states=[board_to_state(b) for b in boards.cpu().numpy()] if N <= 64 else None)

    def close(self):
        for e in self.engines:
            e.close()
