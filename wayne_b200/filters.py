"""Filter placeholders (only the name reaches the FITS header; wayne/filters.py)."""


class F140W(object):
    def __init__(self):
        self.name = 'F140W'
