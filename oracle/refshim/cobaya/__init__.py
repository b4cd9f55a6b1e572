"""Minimal stand-in for the `cobaya` package (absent in this image).

TEST INFRASTRUCTURE ONLY.  It exists so that `oracle/refload.py` can import the
reference's numerical modules from /root/reference *unmodified* when generating
golden vectors in the build container.  Nothing in the product imports it.
"""
