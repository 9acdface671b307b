"""gym.envs.registration.register stand-in; the reference imports it and never calls it."""
registry = {}


def register(id, **kwargs):
    registry[id] = kwargs
