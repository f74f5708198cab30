# SPDX-License-Identifier: Apache-2.0
"""Exception types, same names and hierarchy as ``upkie/exceptions.py``."""


class UpkieException(Exception):
    """Base class for exceptions raised by Upkie agents."""


class FallDetected(UpkieException):
    """Raised when a fall is detected."""


class MissingOptionalDependency(UpkieException):
    """Raised when an optional feature lacks its optional dependency."""


class ModelError(UpkieException):
    """Raised when something is wrong in the robot model."""


class UpkieRuntimeError(UpkieException, RuntimeError):
    """Runtime error, for instance an invalid call to a library function."""
