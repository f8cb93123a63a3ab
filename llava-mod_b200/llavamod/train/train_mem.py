"""``train_mem.py`` -- the entry-point name used by BASELINE.json's north star.  The reference tree has no such file
(its distillation entry points are align_train.py / dpo_train.py); this dispatches to them:
``--stage mimic`` (default) -> align_train.train, ``--stage preference`` -> dpo_train.train."""
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    stage = "mimic"
    if "--stage" in argv:
        i = argv.index("--stage")
        stage = argv[i + 1]
        del argv[i:i + 2]
    if stage in ("mimic", "align", "kd"):
        from .align_train import train
    elif stage in ("preference", "dpo"):
        from .dpo_train import train
    else:
        raise SystemExit("unknown --stage %r (mimic | preference)" % stage)
    return train(argv)


if __name__ == "__main__":
    main()
