#!/usr/bin/env python
"""`python train_mnist.py -g 2 [--syncbn] [--backend nccl|gloo] [--epochs N]` — see
pytorch_distributed_train_b200/cli.py (the framework's counterpart of the reference's run
commands, ref: README.md:97-103)."""
from pytorch_distributed_train_b200.cli import main

if __name__ == "__main__":
    main()
