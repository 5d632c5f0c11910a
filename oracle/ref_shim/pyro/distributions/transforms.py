class Permute:  # imported by reference transforms.py:12 and immediately shadowed by its own class (:174)
    pass
