"""dict with attribute access (reference ``utils/attr.py:4-8``); return type of GaussianParameterize."""


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__.update(kwargs)
