from .base_test_tool import RelDiffComparator, ResultCheck  # noqa: F401
