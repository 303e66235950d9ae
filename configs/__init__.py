"""
The workloads: term tables of the five BASELINE.json configs (plus test tables) as plain dicts
(`specs.py`) and the builder that turns one into a ManagedEnvironment from a given namespace of manager
classes (`env_builder.py`) -- the same `config()` body builds the unmodified reference (under the
oracle's shim) and the CUDA drop-in.  Shared by bench.py, the tests, the oracle and the build tooling;
contains no arithmetic of the path.
"""
