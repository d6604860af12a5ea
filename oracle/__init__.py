"""CPU/PyTorch restatement of the reference hot path (roboticAttack @ a0bef502).

THIS PACKAGE IS TEST INFRASTRUCTURE. Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker / the timed CPU baseline.
Nothing under ``roboticattack_b200/`` imports it; the product path fails loudly without the CUDA extension.

Parity status: the reference has no tests, golden vectors or fixtures of its own for this path (SURVEY.md
section 8c) -- "parity unpinned" by the reference. The oracle is instead pinned against outputs of the
reference's OWN functions executed in the authoring container (``tests/golden/make_golden.py``: the front end
``appply_random_transform.py`` after its one-character indent fix, ``mask_labels`` / ``weighted_loss`` /
``cal_UAD`` of UADA.py, UADA_ddp.py and UPA.py, ``ActionTokenizer``), and against the installed
``transformers`` Llama / cosine schedule for the third-party arithmetic. timm 0.9.10 (vision towers) and
transformers 4.40.1's AdamW are absent from the container: their published algorithms are restated
(``oracle/vit.py``, ``oracle/optim.py``) and remain unpinned.
"""
