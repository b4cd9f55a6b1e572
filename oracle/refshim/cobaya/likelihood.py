from .log import HasLogger


class Likelihood(HasLogger):
    def initialize(self):
        pass

    def initialize_with_provider(self, provider):
        self.provider = provider
