"""Drop-in ``simple_knn`` (the reference does ``from simple_knn._C import distCUDA2``,
geometry/gaussian_base.py:25)."""
