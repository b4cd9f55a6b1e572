"""import stub: `eftpipe/__init__.py` imports `eftpipe.classy`, a subclass of Cobaya's CLASS wrapper (not on this path)"""
from ..theory import Theory


class classy(Theory):
    def initialize(self):
        raise NotImplementedError("CLASS is not available in this image")
