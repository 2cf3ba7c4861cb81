"""LM1B input pipeline — kept here under the reference's module name
(`examples/lm1b/data_utils.py`); the implementation lives in the package."""
from parallax_b200.models.lm1b_data import Dataset, Vocabulary   # noqa: F401
