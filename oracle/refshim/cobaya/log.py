import logging


class LoggedError(Exception):
    """cobaya.log.LoggedError(logger, msg, *args)"""

    def __init__(self, logger=None, *args, **kwargs):
        if isinstance(logger, str):  # tolerate LoggedError("message")
            args, logger = (logger,) + args, None
        msg = (args[0] % args[1:] if len(args) > 1 else args[0]) if args else ""
        super().__init__(msg)


class HasLogger:
    def set_logger(self, lowercase=True, name=None):
        name = name or self.__class__.__name__
        self.log = logging.getLogger(name.lower() if lowercase else name)

    def mpi_info(self, msg, *args):
        self.log.info(msg, *args)

    def mpi_warning(self, msg, *args):
        self.log.warning(msg, *args)

    def mpi_debug(self, msg, *args):
        self.log.debug(msg, *args)


def logger_setup(*args, **kwargs):
    pass
