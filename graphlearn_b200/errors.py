"""Status / error hierarchy (N2).  Same codes and class names as the reference
(graphlearn/python/errors.py:22-215; codes from gRPC plus REQUEST_STOP).
``OutOfRangeError`` is the end-of-epoch signal of every iterator."""
from __future__ import annotations

OK = 0
CANCELLED = 1
UNKNOWN = 2
INVALID_ARGUMENT = 3
DEADLINE_EXCEEDED = 4
NOT_FOUND = 5
ALREADY_EXISTS = 6
PERMISSION_DENIED = 7
RESOURCE_EXHAUSTED = 8
FAILED_PRECONDITION = 9
ABORTED = 10
OUT_OF_RANGE = 11
UNIMPLEMENTED = 12
INTERNAL = 13
UNAVAILABLE = 14
DATA_LOSS = 15
UNAUTHENTICATED = 16
REQUEST_STOP = 17


class GLError(Exception):
    code = UNKNOWN

    def __init__(self, message=""):
        super().__init__(message)
        self.message = message

    @property
    def error_code(self):
        return self.code


OpError = GLError
BaseError = GLError          # the reference's name for the base class (python/errors.py:22)


def _mk(name, code, doc):
    return type(name, (GLError,), {"code": code, "__doc__": doc})


CancelledError = _mk("CancelledError", CANCELLED, "operation was cancelled")
UnknownError = _mk("UnknownError", UNKNOWN, "unknown error")
InvalidArgumentError = _mk("InvalidArgumentError", INVALID_ARGUMENT, "invalid argument")
DeadlineExceededError = _mk("DeadlineExceededError", DEADLINE_EXCEEDED, "deadline exceeded")
NotFoundError = _mk("NotFoundError", NOT_FOUND, "entity not found")
AlreadyExistsError = _mk("AlreadyExistsError", ALREADY_EXISTS, "entity already exists")
PermissionDeniedError = _mk("PermissionDeniedError", PERMISSION_DENIED, "permission denied")
UnauthenticatedError = _mk("UnauthenticatedError", UNAUTHENTICATED, "unauthenticated")
ResourceExhaustedError = _mk("ResourceExhaustedError", RESOURCE_EXHAUSTED, "resource exhausted")
FailedPreconditionError = _mk("FailedPreconditionError", FAILED_PRECONDITION, "failed precondition")
AbortedError = _mk("AbortedError", ABORTED, "aborted")
OutOfRangeError = _mk("OutOfRangeError", OUT_OF_RANGE, "iteration reached the end of an epoch")
UnimplementedError = _mk("UnimplementedError", UNIMPLEMENTED, "not implemented")
InternalError = _mk("InternalError", INTERNAL, "internal error")
UnavailableError = _mk("UnavailableError", UNAVAILABLE, "service unavailable")
DataLossError = _mk("DataLossError", DATA_LOSS, "data loss")
RequestStopError = _mk("RequestStopError", REQUEST_STOP, "stop requested")

_CODE_TO_EXC = {c.code: c for c in [
    CancelledError, UnknownError, InvalidArgumentError, DeadlineExceededError, NotFoundError,
    AlreadyExistsError, PermissionDeniedError, UnauthenticatedError, ResourceExhaustedError,
    FailedPreconditionError, AbortedError, OutOfRangeError, UnimplementedError, InternalError,
    UnavailableError, DataLossError, RequestStopError]}


def exception_type_from_error_code(code):
    return _CODE_TO_EXC[code]


def error_code_from_exception_type(cls):
    return cls.code


def raise_exception_on_not_ok_status(code, message=""):
    if code != OK:
        raise _CODE_TO_EXC.get(code, UnknownError)(message)
