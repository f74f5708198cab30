# SPDX-License-Identifier: Apache-2.0
"""Error types of the package.

The NAMES and the inheritance follow the reference (``upkie/exceptions.py``) so that ``except upkie.exceptions.X``
blocks written for it keep catching the same situations here; when the reference package is importable its classes
are re-used outright, which makes ``isinstance`` checks interchangeable between the two packages.
"""

try:  # pragma: no cover - the reference is not installed in this image
    from upkie.exceptions import (  # type: ignore  # noqa: F401
        FallDetected,
        MissingOptionalDependency,
        ModelError,
        PerformanceIssue,
        SpineError,
        UpkieException,
        UpkieRuntimeError,
        UpkieTimeoutError,
    )
except ImportError:

    class UpkieException(Exception):
        """Root of every error this package raises on purpose."""

    def _derive(name: str, doc: str, *extra_bases: type) -> type:
        return type(name, (UpkieException,) + extra_bases, {"__doc__": doc, "__module__": __name__})

    FallDetected = _derive("FallDetected", "The base pitch went past the fall threshold.")
    MissingOptionalDependency = _derive(
        "MissingOptionalDependency", "A feature was requested whose optional third-party package is not installed.")
    ModelError = _derive("ModelError", "The robot description is inconsistent or not supported by the kernels.")
    PerformanceIssue = _derive("PerformanceIssue", "A configuration that would silently run slowly was detected.")
    SpineError = _derive("SpineError", "A spine reported an error through its mailbox.")
    UpkieRuntimeError = _derive(
        "UpkieRuntimeError", "A library call failed or was used with invalid arguments; also a RuntimeError.", RuntimeError)
    UpkieTimeoutError = _derive("UpkieTimeoutError", "A peer did not answer in time; also a TimeoutError.", TimeoutError)
