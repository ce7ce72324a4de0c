def _filter_items_from_sparse_matrix(items, query_items):
    """Restrict the CSR filter to the whitelisted columns, renumbered to whitelist positions."""
    return query_items[:, items]
