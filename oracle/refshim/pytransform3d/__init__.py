"""Throw-away stand-in for the un-vendored `pytransform3d` dependency.

TEST INFRASTRUCTURE ONLY: lets `/root/reference/distance3d` be imported in
the build container (gen_golden.py, validate_oracle.py).  Never imported by the
product package.
"""
