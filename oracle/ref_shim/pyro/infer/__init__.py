SVI = None  # imported by reference transforms.py:13, never used
