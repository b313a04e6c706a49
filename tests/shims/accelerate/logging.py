"""TEST SHIM: accelerate.logging.get_logger — a LoggerAdapter whose calls accept `main_process_only` (default True)
and `in_order` (train.py:35,123; evaluate.py:32)."""
import logging

_STATE = {"ready": False, "main": True}


class MultiProcessAdapter(logging.LoggerAdapter):
    def log(self, level, msg, *args, **kwargs):
        main_only = kwargs.pop("main_process_only", True)
        kwargs.pop("in_order", None)
        if self.isEnabledFor(level) and (_STATE["main"] or not main_only):
            kwargs.setdefault("stacklevel", 2)
            msg, kwargs = self.process(msg, kwargs)
            self.logger.log(level, msg, *args, **kwargs)


def get_logger(name: str, log_level: str = None):
    logger = logging.getLogger(name)
    if log_level is not None:
        logger.setLevel(log_level.upper())
        logger.root.setLevel(log_level.upper())
    return MultiProcessAdapter(logger, {})
