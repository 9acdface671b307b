"""Stand-in for the `gym` package (TEST INFRASTRUCTURE ONLY).

`gym>=0.26.2` (reference setup.py:14) is not installed in this image and there is no
network.  The unmodified reference imports exactly three things from it:

  * gym.envs.registration.register        (gym_soccer/__init__.py:1, never called)
  * gym.spaces.Dict / gym.spaces.Discrete (soccer_simultaneous_env.py:3, :126-131)
  * gym.envs.toy_text.utils.categorical_sample (soccer_simultaneous_env.py:2, :395, :414)

This package restates the published gym 0.26.2 behaviour of those three, nothing else.
It is placed ahead of /root/reference on sys.path by oracle/ref_harness.py so that the
reference can be imported and driven, unmodified, to produce golden vectors.
"""
from . import spaces  # noqa: F401
