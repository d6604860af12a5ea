"""CPU/PyTorch restatement of the reference hot path (roboticAttack @ a0bef502).

THIS PACKAGE IS TEST INFRASTRUCTURE. Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker / the timed CPU baseline.
Nothing under ``roboticattack_b200/`` imports it; the product path fails loudly without the CUDA extension.

Parity status: the reference has no tests, golden vectors or fixtures of its own for this path (SURVEY.md
section 8c) -- "parity unpinned" by the reference. The oracle is instead pinned against outputs of the
reference's OWN functions executed in the authoring container (``tests/golden/make_golden.py``: the front end
``appply_random_transform.py`` after its one-character indent fix, ``mask_labels`` / ``weighted_loss`` /
``cal_UAD`` of UADA.py, UADA_ddp.py and UPA.py, ``ActionTokenizer``), against the reference's own model
class and attack loops run on the CPU (``make_golden_glue.py``: ``OpenVLAForActionPrediction.forward`` / ``predict_action``;
``make_golden_loop.py``: ``patchattack_unconstrained`` of UADA.py / UPA.py / TMA.py with their validation passes), and against
the installed ``transformers`` Llama / cosine schedule for the third-party arithmetic. timm 0.9.10 (vision towers) is absent
from the container: its published algorithm is restated in ``oracle/vit.py`` and cross-checked against transformers'
DINOv2-with-registers and SigLIP vision models on shared random weights. transformers 4.40.1's AdamW (removed upstream) is
restated in ``oracle/optim.py`` and checked against its definition only -- the one piece that remains unpinned.
"""
