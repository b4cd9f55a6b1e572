import logging


class LoggedError(Exception):
    def __init__(self, logger, *args, **kwargs):
        msg = args[0] % args[1:] if len(args) > 1 else (args[0] if args else "")
        super().__init__(msg)


class HasLogger:
    def set_logger(self, lowercase=True, name=None):
        self.log = logging.getLogger(name or self.__class__.__name__)

    def mpi_info(self, msg, *args):
        pass

    def mpi_warning(self, msg, *args):
        pass

    def mpi_debug(self, msg, *args):
        pass


def logger_setup(*args, **kwargs):
    pass
