"""``python -m hallucidet_b200.run <script.py> [script args...]`` -- run one of the reference's scripts
(train_hallucidet.py, eval_hallucidet.py; launched from the reference's repository root, as its README says) with the
B200 hot path bound in (``hallucidet_b200.patch.apply``), without editing the reference.

Options before the script name:
  --train-detector   leave the detector trainable and on torchvision's backbone (the frozen dgrad-only backbone is only
                     valid when ``Config.Detector.train_det`` is False, the HalluciDet setting)
  --stock-tail       keep the reference's own eval_forward_* (per-image loops) instead of the batched restatement
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    freeze, fast_tail = True, True
    while argv and argv[0].startswith("--"):
        flag = argv.pop(0)
        if flag == "--train-detector":
            freeze = False
        elif flag == "--stock-tail":
            fast_tail = False
        else:
            raise SystemExit(f"hallucidet_b200.run: unknown option {flag}")
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0]
    here = os.path.dirname(os.path.abspath(script))
    for p in (os.getcwd(), here):                              # the scripts import ``src.*`` relative to the checkout root
        if p not in sys.path:
            sys.path.insert(0, p)
    from . import patch
    bound = patch.apply(freeze_detector=freeze, fast_tail=fast_tail)
    print(f"[hallucidet_b200] bound: {', '.join(bound)}", file=sys.stderr)
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
