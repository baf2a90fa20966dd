"""TEST INFRASTRUCTURE ONLY — CPU oracle for the cross-attention ViT hot path.

Nothing under ``oracle/`` is part of the product. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline. The product
(``cross-attention-vit_b200/cavit``) never imports this package and has no CPU
fallback: it fails loudly when ``libcavit_sm100a.so`` is missing.

Contents
--------
functional.py   independent torch restatement (fp32/fp64) of ``ModelCross.forward``
                (/root/reference/model_cross.py:11-212) and ``ModelVIT.forward``
                (/root/reference/modelv3.py:18-147), written from SURVEY.md §3.2 / §A.
weights.py      deterministic weight / input construction shared by fixtures and tests.
ref_loader.py   imports the UNMODIFIED reference from /root/reference with stub modules
                for its absent third-party imports (only usable in the build container).
encoders.py     restatement of the CNN-stem models ``ViT`` (model.py) and ``ViT3D`` (modelv2.py).
gen_golden.py   generates tests/golden/*.pt by running the real reference (committed
                together with the vectors it made).

staging.py      numpy restatement of the un-augmented input chain of dataset_ucsf.py (nibabel read
                scaling + MONAI ResizeWithPadOrCrop); PARITY UNPINNED (both packages absent), see its header.
metrics.py      restatement of log_stats / compute_metrics (utils.py, model_cross.py:243-255); pinned against
                scikit-learn (torchmetrics absent: PARITY UNPINNED against it).

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself run in the build container
(tests/golden/, produced by gen_golden.py) and, when /root/reference is present,
directly against the imported reference modules (tests/test_oracle_vs_reference.py).
"""
