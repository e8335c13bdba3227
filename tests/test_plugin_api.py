"""Plugin API: the stand-alone Parameter / Benchmark / REGISTRY behave like the reference's
(stencil_benchmarks/benchmark.py:50-220).  Same scenarios as the reference's own unit test
(stencil_benchmarks/test/test_benchmark.py:62-112), written against this package."""

import pytest

from stencil_benchmarks_b200 import benchmark


class Sample(benchmark.Benchmark):
    flag = benchmark.Parameter("a switch", False)
    count = benchmark.Parameter("a count", 1)

    def setup(self):
        self.half = self.count // 2

    def run(self):
        return self.half


class Child(Sample):
    label = benchmark.Parameter("a label", "0")
    count = benchmark.Parameter("a count with another default", 7)

    def setup(self):
        super().setup()
        self.total = int(self.flag) + self.count + int(self.label)

    def run(self):
        return self.total


def test_registration_and_parameter_collection():
    assert Sample in benchmark.REGISTRY and Child in benchmark.REGISTRY
    assert Sample.parameters == {"flag": benchmark.Parameter("a switch", False),
                                 "count": benchmark.Parameter("a count", 1)}
    # own declarations override inherited ones (benchmark.py:147-158)
    assert Child.parameters["count"].default == 7
    assert set(Child.parameters) == {"flag", "count", "label"}


def test_abstract_classes_are_not_registered():
    class Abstract(benchmark.Benchmark):
        pass

    assert Abstract not in benchmark.REGISTRY


def test_init_validates_and_calls_setup():
    b = Sample(flag=True, count=42)
    assert (b.flag, b.count, b.half) == (True, 42, 21)
    assert b.parameters == {"flag": True, "count": 42}
    assert b.run() == 21 and b() == 21
    with pytest.raises(benchmark.ParameterError, match='invalid value for argument "flag"'):
        Sample(flag=3, count=42)
    with pytest.raises(ValueError, match="unsupported arguments"):
        Sample(nonsense=1)
    assert Child().total == 7


def test_reassignment_revalidates():
    b = Child(flag=True, count=42, label="5")
    b.flag = False
    assert b.flag is False and b.parameters["flag"] is False
    with pytest.raises(benchmark.ParameterError):
        b.flag = 3


def test_parameter_rules():
    P = benchmark.Parameter
    with pytest.raises(ValueError):
        P("no default, no type")
    with pytest.raises(ValueError):
        P("mixed tuple", (1, 2.0))
    with pytest.raises(ValueError):
        P("empty", ())
    with pytest.raises(ValueError):
        P("wrong dtype", 1, dtype=str)
    with pytest.raises(ValueError):
        P("wrong nargs", (1, 2), nargs=3)
    required = P("required", dtype=str, nargs=1)
    with pytest.raises(benchmark.ParameterError, match="value is required"):
        required.validate(None)
    triple = P("triple", (1, 2, 3))
    assert triple.nargs == 3 and triple.dtype is int
    assert triple.validate([4, 5, 6]) == [4, 5, 6]
    for bad in (4, (1, 2), (1, 2, "3")):
        with pytest.raises(benchmark.ParameterError):
            triple.validate(bad)
    choice = P("choice", "a", choices=["a", "b"])
    assert choice.validate(None) == "a"
    with pytest.raises(benchmark.ParameterError, match="choices are"):
        choice.validate("c")


def test_parameters_attribute_is_reserved():
    with pytest.raises(AttributeError):
        class Bad(benchmark.Benchmark):  # noqa: F841
            parameters = 3
