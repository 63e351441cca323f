"""CPU oracle for the appearance-flow hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy (and, for multi-threaded timing, torch-CPU), the
algorithm of the reference's novel-view-synthesis training path:
``dyn_mult_view/mv3d/utils/tf_utils.py:18-98`` and the graphs in
``dyn_mult_view/multi_view_model/*.py``, plus the TensorFlow-1.3 semantics those
lines delegate to (SAME padding, conv2d_transpose, tf.contrib.resampler, Adam).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it -- as the checker or the reported CPU
baseline, never as the product.  Nothing under ``dynamic_multiview_3d_b200/``
imports this package.

PARITY UNPINNED: the reference needs Python 2.7 + TensorFlow 1.3,
neither of which exists in this image, and its only test for this path
(``multi_view_model/tests/test_resampler.py``) asserts nothing.  The restatement is
pinned instead by hand-derived known answers, float64 finite differences, a
restatement of test_resampler.py on its rectangle fixture, and agreement with
``torch.nn.functional.grid_sample`` / ``torch.nn.functional.conv2d`` on CPU.
"""
