"""CPU/fp32 oracle for the HalluciDet hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``hallucidet_b200`` (the product)
never imports it and fails loudly when its CUDA library is missing.

What it is: a plain-PyTorch fp32 restatement (``torch.nn.functional`` calls on state-dict
tensors, no ``nn.Module`` tree of the reference) of the reference's algorithm on the path
SURVEY.md section 8 names:

    oracle/unet.py        smp.Unet('resnet34') forward (train / eval BN) + reference initialisers
    oracle/backbone.py    torchvision ResNet-50-FPN backbone forward with frozen BN (Faster R-CNN and RetinaNet variants)
    oracle/transform.py   CustomGeneralizedRCNNTransform: identity normalise + nearest resize + batch + box rescale
    oracle/losses.py      pixel regulariser (MSE / L1) and the loss assembly of train_hallucidet.py:161-209
    oracle/detector.py    eval_forward_fasterrcnn / eval_forward_retinanet (losses in eval mode) over torchvision heads
    oracle/step.py        the assembled train step (U-Net -> regulariser + detector loss -> backward)

Every function cites the reference file:line it follows (paths relative to the reference
repo root; ``TV:`` = the installed torchvision, third-party code that the reference itself
calls).

Parity pin: the reference ships no tests, golden vectors or fixtures, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, imported in the build
container by ``tests/golden/make_golden.py`` (stub recipe of SURVEY.md section 8c) and committed
as small fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.
"""
