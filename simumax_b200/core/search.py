"""Search API (filled in below): micro-batch / recompute / parallel-strategy search."""


class SearchMixin:
    pass
