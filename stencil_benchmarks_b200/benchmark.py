"""Plugin API of the B200 backend: ``Parameter`` / ``Benchmark`` / ``REGISTRY``.

When the reference package ``stencil_benchmarks`` is importable, its own
classes are re-exported, so every benchmark defined in this package is a
subclass of the reference's ``Benchmark``, lands in the reference's
``REGISTRY`` through its metaclass and shows up in ``sbench`` as
``sbench stencils b200 ...`` / ``sbench stream b200 ...``
(stencil_benchmarks/benchmark.py:140-164, stencil_benchmarks/cli.py:47-50).

When it is not importable (e.g. on a GPU box that only has this repository),
the stand-alone implementation below provides the same behaviour
(stencil_benchmarks/benchmark.py:50-220):

* ``Parameter(description, default, dtype, nargs, choices)`` with the same
  consistency checks and ``validate`` rules;
* a metaclass that gathers ``Parameter`` class attributes (own ones override
  inherited ones), forbids an attribute called ``parameters`` and registers
  every non-abstract class in ``REGISTRY``;
* ``Benchmark(**kwargs)``: unknown arguments -> ``ValueError``, invalid values ->
  ``ParameterError('invalid value for argument "x": ...')``, then ``setup()``;
  parameter values are readable/writable as attributes (writes re-validate).
"""

try:  # reference available: plug into it
    from stencil_benchmarks.benchmark import (  # noqa: F401
        REGISTRY,
        Benchmark,
        BenchmarkMeta,
        ExecutionError,
        Parameter,
        ParameterError,
    )

    HAVE_REFERENCE = True
except ImportError:
    HAVE_REFERENCE = False

    import abc
    import inspect

    REGISTRY = set()

    class ParameterError(ValueError):
        pass

    class ExecutionError(RuntimeError):
        pass

    def _infer(default):
        """(element type, arity) implied by a default value."""
        many = isinstance(default, (tuple, list))
        values = list(default) if many else [default]
        if not values:
            raise ValueError("can not use empty tuple as default")
        kinds = {type(v) for v in values}
        if len(kinds) != 1:
            raise ValueError("different types in default tuple")
        return kinds.pop(), (len(values) if many else 1)

    class Parameter:
        """Typed, documented benchmark option (becomes a CLI flag in ``sbench``)."""

        def __init__(self, description, default=None, dtype=None, nargs=None, choices=None):
            if default is None:
                if dtype is None or nargs is None:
                    raise ValueError("dtype and nargs must be given if default is None")
            else:
                implied_dtype, implied_nargs = _infer(default)
                if dtype not in (None, implied_dtype):
                    raise ValueError("iconsistent default and dtype values")
                if nargs not in (None, implied_nargs):
                    raise ValueError("inconsistent default and nargs values")
                dtype, nargs = implied_dtype, implied_nargs
            self.description, self.default = description, default
            self.dtype, self.nargs = dtype, nargs
            self.choices = None if choices is None else tuple(choices)

        def _typed(self, value):
            if self.nargs == 1:
                if isinstance(value, self.dtype):
                    return
                raise ParameterError(
                    f'wrong type of argument "{value}", '
                    f' expected one of type "{self.dtype.__name__}"')
            if not isinstance(value, (tuple, list)):
                raise ParameterError(
                    f'{self.nargs} arguments of type "{self.dtype.__name__}" required, found "{value}"')
            if len(value) != self.nargs:
                raise ParameterError(f'wrong number of arguments in argument "{value}"')
            if any(not isinstance(v, self.dtype) for v in value):
                raise ParameterError(f'wrong type in argument "{value}"')

        def validate(self, value):
            """The value to use for `value` (None selects the default); raises ParameterError."""
            if value is None:
                value = self.default
                if value is None:
                    raise ParameterError("value is required")
            self._typed(value)
            if self.choices is not None and value not in self.choices:
                listed = ", ".join(f'"{c}"' for c in self.choices)
                raise ParameterError(f'unsupported argument value "{value}", choices are {listed}')
            return value

        def __repr__(self):
            return (
                f"{type(self).__name__}(description={self.description!r}, dtype={self.dtype!r}, "
                f"nargs={self.nargs!r}, default={self.default!r})"
            )

        def __eq__(self, other):
            mine = (self.description, self.dtype, self.nargs, self.default)
            return mine == (other.description, other.dtype, other.nargs, other.default)

        __hash__ = None

    class BenchmarkMeta(abc.ABCMeta):
        def __new__(mcs, name, bases, namespace):
            if "parameters" in namespace:
                raise AttributeError("Benchmark classes must not define an attribute `parameters`")
            collected = {}
            for base in bases:
                collected.update(getattr(base, "parameters", {}))
            plain = {}
            for key, value in namespace.items():
                if isinstance(value, Parameter):
                    collected[key] = value
                else:
                    plain[key] = value
            plain["parameters"] = collected
            cls = super().__new__(mcs, name, bases, plain)
            if not inspect.isabstract(cls):
                REGISTRY.add(cls)
            return cls

    class Benchmark(metaclass=BenchmarkMeta):
        def __init__(self, **kwargs):
            unknown = set(kwargs) - set(self.parameters)
            if unknown:
                raise ValueError(
                    "unsupported arguments " + ", ".join(f'"{arg}"' for arg in unknown)
                )
            values = {}
            for arg, param in self.parameters.items():
                try:
                    values[arg] = param.validate(kwargs.get(arg))
                except ParameterError as error:
                    raise ParameterError(
                        f'invalid value for argument "{arg}": ' + error.args[0]
                    ) from None
            # from here on `parameters` on the instance holds the values
            self.parameters = values
            self.setup()

        def __setattr__(self, name, value):
            if name in self.parameters:
                self.parameters[name] = type(self).parameters[name].validate(value)
            else:
                super().__setattr__(name, value)

        def __getattr__(self, name):
            if name in self.parameters:
                return self.parameters[name]
            return super().__getattribute__(name)

        def setup(self):
            """Prepare the benchmark (allocate, compile, load)."""

        @abc.abstractmethod
        def run(self):
            """Run once; return a result dict (or a list of them)."""

        def __call__(self):
            return self.run()

        def __repr__(self):
            args = ", ".join(f"{k}={v!r}" for k, v in self.parameters.items())
            return f"{type(self).__name__}({args})"
