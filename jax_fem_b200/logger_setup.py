"""Logger named like the reference's (jax_fem/logger_setup.py:5-29)."""
import logging


def setup_logger(name='jax_fem_b200', level=logging.INFO):
    logger = logging.getLogger(name)
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter('[%(asctime)s][%(levelname)s][%(name)s] %(message)s', '%m-%d %H:%M:%S'))
        logger.addHandler(handler)
        logger.setLevel(level)
        logger.propagate = False
    return logger
