"""Stand-in for the four MONAI symbols `/root/reference/model` imports.

Test infrastructure only: it lets `tests/golden/make_golden.py` import the UNMODIFIED reference
model in the build container, where monai==1.5.0 is not installed.  Never imported by the product.
"""
