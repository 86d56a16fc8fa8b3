import gym


def register(id, entry_point, **kwargs):
    gym._reg[id] = entry_point
