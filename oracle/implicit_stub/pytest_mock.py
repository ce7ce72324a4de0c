"""Name-only shim so the reference's test modules that `from pytest_mock import MockerFixture` can be collected here
(pytest-mock is a reference dev dependency that is not installed; test infrastructure only)."""


class MockerFixture:  # pragma: no cover
    pass
