"""polycracker_b200: B200-native hot path of polyCRACKER (k-mer count -> repeat selection -> occurrence matrix ->
TF-IDF).  In-tree this directory is imported through the `polycracker_b200/` alias package (a hyphenated directory
name is not importable); installed, setup.py maps it to the same package name."""
__version__ = "0.1.0"
