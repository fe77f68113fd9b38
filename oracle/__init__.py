"""Test infrastructure: CPU restatement of the reference's retrieval-ranking path (see
reference_path.py).  Never imported by the product package."""
