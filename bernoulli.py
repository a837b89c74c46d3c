"""`import bernoulli` drop-in (reference: moira/bernoullimodule.c; imported at moira/moira.py:247-251
and moira/test/test_moira.py:7-12).  Re-exports the CUDA-backed shim."""
from moira_b200.bernoulli import calculate_errors_PB  # noqa: F401
from moira_b200.bernoulli import __doc__  # noqa: F401
