"""Logging helper (reference: python/mxnet/log.py)."""
import logging
import sys

CRITICAL, ERROR, WARNING, INFO, DEBUG, NOTSET = (logging.CRITICAL, logging.ERROR, logging.WARNING, logging.INFO,
                                                 logging.DEBUG, logging.NOTSET)


def get_logger(name=None, filename=None, filemode=None, level=WARNING):
    logger = logging.getLogger(name)
    if name is not None and not getattr(logger, "_init_done", None):
        logger._init_done = True
        hdlr = logging.FileHandler(filename, filemode or "a") if filename else logging.StreamHandler(sys.stderr)
        hdlr.setFormatter(logging.Formatter("%(asctime)s %(levelname).1s %(process)d %(filename)s:%(lineno)d] %(message)s"))
        logger.addHandler(hdlr)
        logger.setLevel(level)
    return logger


getLogger = get_logger
