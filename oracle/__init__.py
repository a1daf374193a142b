"""CPU oracle for the DeepMod `detect` hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``deepmod_b200/`` imports this package.  The only permitted
importers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and only as the checker or as the
timed CPU baseline -- never as the thing shipped.

Parity status: the reference has no tests, golden vectors or fixtures for this
path, and its arithmetic lives in TensorFlow 1.x
(un-vendored, not installable here; every shipped ``.meta`` records producer
1.8.0).  The oracle is pinned instead against

* the reference's own python for window assembly / batching / label write-back
  (``bin/DeepMod_scripts/myDetect.py:787-903``), imported unmodified in the
  build container with ``tensorflow``/``h5py`` stubbed (``oracle/ref_harness.py``)
  -- the committed fixtures under ``tests/golden/`` were generated through it by
  ``tests/golden/make_golden.py``;
* the frozen GraphDefs under ``train_deepmod/rnn_*/*.meta`` for the op order of
  the BiLSTM restatement (``tests/test_oracle_graph.py``).

Modules: ``detect_ref`` / ``bilstm`` / ``tf_bundle`` (the core path), ``align_ref`` (SAM/CIGAR walk, pinned by
running the unmodified ``handle_line`` + ``handle_record``), ``cluster_ref`` (CpG-cluster second pass, pinned by
running the unmodified ``hm_cluster_predict.py`` / ``sum_chr_mod.py``), ``signal_ref`` (raw-signal normalisation,
pinned by the unmodified ``mnormalized``), ``cells`` (model of the packed accumulator), ``ref_harness``.

The TensorFlow arithmetic itself (Eigen GEMM summation order, Eigen's float
sigmoid/tanh) cannot be executed here: **TF-level parity is unpinned**; the
restatement follows the graph op-for-op in fp64/fp32.
"""
