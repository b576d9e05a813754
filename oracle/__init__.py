"""CPU oracle for the vbt hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the timed CPU arm.  The product
(`vbt_b200`) never imports this package and fails loudly when its CUDA library is
missing.

Modules (each function cites the reference file:line it restates):

* ``velocity``    -- RunningAverage / VelocityTracker / Phase state machine and the
                     pandas smoothing of plot.py (pinned: live reference classes,
                     pandas 3.0.2, 34/34 figs_ocsort labels).
* ``ocsort``      -- OC-SORT as configured at track.py:157 ([3P-MEM]; Kalman filter
                     pinned bit-exactly by dfs_ocsort replay, association internals
                     "parity unpinned").
* ``resize``      -- tf.image.resize bilinear + truncating cast (odt.py:15-18)
                     ([3P-MEM]; parity unpinned, no tensorflow in the container).
* ``postprocess`` -- TFLite_Detection_PostProcess fast-NMS ([3P-MEM]; weakly pinned by
                     dfs/eval_detections.pkl.gz: 25 outputs, 1/256 score grid).
* ``effdet``      -- int8 EfficientDet-Lite0/1/2 graph ([3P-MEM]; MAC counts pinned to
                     models/*.log:110; numerics parity unpinned -- the .tflite blobs
                     are absent from the reference checkout).
"""
