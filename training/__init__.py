"""Shim package: resolves the reference's ``training.coarse.eval_epoch`` import (``evaluation/pipeline.py:28``) to the
B200-native drop-in.  The training loops themselves are out of scope."""
