"""gym 0.26.2 `gym/envs/toy_text/utils.py::categorical_sample`, restated.

Published behaviour: cumulative sum of the probabilities, ONE uniform draw from
`np_random.random()`, strict `>` comparison, `argmax` of the boolean vector (so an
all-False vector yields index 0).
"""
import numpy as np


def categorical_sample(prob_n, np_random):
    prob_n = np.asarray(prob_n)
    csprob_n = np.cumsum(prob_n)
    return np.argmax(csprob_n > np_random.random())
