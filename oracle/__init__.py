"""ORACLE — test infrastructure only (see oracle/head.py header).  Not part of the product."""
