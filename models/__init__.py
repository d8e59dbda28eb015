"""Dotted-path shims: ``models.*`` of the reference resolves to the B200-native modules, so that
``evaluation.pipeline`` / ``training.coarse.eval_epoch`` (and ``torch.load`` of whole-module pickles) find the
classes where they expect them."""
