"""CPU oracle for the MPNTrackSeg message-passing hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or the timed CPU baseline.  Nothing
under ``mpntrackseg_b200/`` imports it; the product path fails loudly when the CUDA
library is missing.

It restates, in plain CPU PyTorch (fp32, the arithmetic the reference itself runs on
CPU), the algorithms of

* ``src/mot_neural_solver/utils/graph.py:6-124``          (time-valid pairs, KNN mask, edge features)
* ``src/mot_neural_solver/data/mot_graph.py:195-221,283-316`` (edge construction / graph assembly)
* ``src/mot_neural_solver/models/mpn.py:33-394`` + ``models/mlp.py:4-28`` + ``models/cnn.py:4-84``
* ``src/mot_neural_solver/tracker/mpn_tracker.py:96-141``  (prune -> forward -> sigmoid -> scatter back)
* ``src/mot_neural_solver/tracker/mpn_tracker.py:143-210`` + ``utils/graph.py:165-207`` (sliding windows over a
  sequence, per-edge averaging, undirected merge, pruning at 0.5; ``tracker_ref.py``)
* ``src/mot_neural_solver/data/mot_graph.py:223-262``      (edge labels of the network-flow formulation)
* ``src/mot_neural_solver/pl_module/pl_module.py:88-120``  (weighted BCE loss)
* ``src/mot_neural_solver/utils/evaluation.py:370-414``, ``tracker/projectors.py:11-67``,
  ``tracker/mpn_tracker.py:231-248``  (constraint statistics, greedy rounding, identities; ``rounding_ref.py``, numpy)

and of the third-party ``torch-scatter==2.0.4`` calls made on that path
(``scatter_add`` = index-add with zero fill, ``scatter_softmax`` =
``exp(x - segmax) / (segsum + 1e-12)``), which is not vendored in the reference.

Parity pinning: the reference holds NO test, golden vector or fixture for this path
(SURVEY.md section 4 / 8c).  The oracle is therefore pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports the unmodified reference
modules from ``/root/reference/src`` (with a ``torch_scatter`` stand-in), runs them on
seeded synthetic windows and commits the outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this oracle against those files (model cases, the
configs[4] window with 4,500 nodes, a 5,400-node window, tracker window + sequence, edge labels,
rounding + identities).
"""
