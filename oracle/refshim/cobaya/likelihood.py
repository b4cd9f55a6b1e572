from .theory import Theory


class Likelihood(Theory):
    """cobaya.likelihood.Likelihood: a Theory whose `calculate` stores `state["logp"]`"""

    type = []

    def logp(self, **params_values):
        return None

    def calculate(self, state, want_derived=True, **params_values_dict):
        derived = {} if want_derived else None
        state["logp"] = self.logp(_derived=derived, **params_values_dict)
        if derived is not None:
            state["derived"].update(derived)

    @property
    def current_logp(self):
        return self.current_state["logp"]
