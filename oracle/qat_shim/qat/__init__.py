"""TEST INFRASTRUCTURE ONLY -- pure-Python stand-in for the (closed, absent)
myQLM ``qat`` package, just large enough for the unmodified reference modules
to import and run.  Never imported by the product package ``openvqe_b200``."""
__shim__ = True
