"""CPU oracle for the alphaFive self-play hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU arm.  The product path (``alphafive_b200``) never
imports this package and fails loudly when its CUDA library is missing.

Contents (each function cites the reference file:line it restates):

* ``oracle.rules``  -- numpy restatement of the rule/encoding functions of the
  reference's ``utils.py:156-296``.
* ``oracle.mcts``   -- array-based restatement of ``genData/player.py`` (the
  dict-keyed transposition-table MCTS), reproducing the reference's dtype chain.
* ``oracle.net``    -- torch-CPU fp32 restatement of ``genData/network.py:52-97``
  (the TF1 graph cannot be imported: TensorFlow is absent and unpinned).
* ``oracle.make_golden`` -- imports the UNMODIFIED reference from
  ``/root/reference`` (build container only) and writes ``tests/golden/*.npz``.

Parity status: the reference ships no tests, so the pins are (i) golden vectors
produced by running the reference's own ``utils`` / ``Player`` code here
(``tests/golden/rules_*.npz``, ``mcts_kat_*.npz``), (ii) invariants of the
shipped replay buffer ``data_buffer/data6960.pkl`` (``replay_sample.npz``) and
(iii) the logged losses at ckpt-6960 for the network restatement
(``ckpt6960.npz``) and (iv) the human-vs-AI game the reference ships as
``tmp/five_6960.gif`` (GUI.py:184-186; decoded into ``gui_game_6960.npz``): the 29
moves its AI -- ckpt-6960 through the TensorFlow net, Player(training=False),
542 / 642 simulations -- played are reproduced move for move by ``oracle.net``
driving ``oracle.mcts`` (tests/test_oracle_gui_game.py), close calls included
(198 against 193 visits).  Raw NN output vectors do not exist (TensorFlow cannot
run here); the network restatement is pinned through (iii) and, end to end,
through (iv).
"""
