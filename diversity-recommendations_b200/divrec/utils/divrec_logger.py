"""stdlib logging helpers with the reference's names (divrec/utils/divrec_logger.py:4-27)."""
import logging
from typing import Optional

LOG_FORMAT = "%(asctime)s %(name)s [%(levelname)s] %(message)s"


def _configured(handler: logging.Handler) -> logging.Handler:
    handler.setLevel(logging.INFO)
    handler.setFormatter(logging.Formatter(LOG_FORMAT))
    return handler


def get_file_handler(filepath: str) -> logging.Handler:
    return _configured(logging.FileHandler(filepath))


def get_stream_handler() -> logging.Handler:
    return _configured(logging.StreamHandler())


def get_logger(name: str, filepath: Optional[str] = None) -> logging.Logger:
    log = logging.getLogger(name)
    log.setLevel(logging.INFO)
    if filepath is not None:
        log.addHandler(get_file_handler(filepath))
    log.addHandler(get_stream_handler())
    return log
