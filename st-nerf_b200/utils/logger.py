"""utils/logger.py:12-30 of the reference: `setup_logger(name, save_dir, distributed_rank)` (imported by demo/*.py)."""
import logging
import os
import sys


def setup_logger(name, save_dir, distributed_rank):
    log = logging.getLogger(name)
    log.setLevel(logging.DEBUG)
    if distributed_rank > 0:             # only the master process gets handlers
        return log
    fmt = logging.Formatter("%(asctime)s %(name)s %(levelname)s: %(message)s")
    handlers = [logging.StreamHandler(stream=sys.stdout)]
    if save_dir:
        handlers.append(logging.FileHandler(os.path.join(save_dir, "log.txt"), mode="w"))
    for h in handlers:
        h.setLevel(logging.DEBUG)
        h.setFormatter(fmt)
        log.addHandler(h)
    return log
