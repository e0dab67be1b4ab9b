from . import Add, Concatenate, Multiply, add, concatenate, multiply  # noqa: F401
