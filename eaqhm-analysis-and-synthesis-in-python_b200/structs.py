"""Result records with the reference's attribute names (structs.py:7-34)."""


class Deterministic:
    """Per-frame output record of the drop-in (reference: structs.py:7-22).

    Like the reference, the constructor stores its ``amplitudes`` argument as ``.ak``
    while the packing step assigns ``.amplitudes`` afterwards (quirk Q8).
    """

    def __init__(self, ti=[], isSpeech=False, isVoiced=False, a0=[], amplitudes=[], frange=[], pk=[]):
        self.ti = ti
        self.isSpeech = isSpeech
        self.isVoiced = isVoiced
        self.a0 = a0
        self.ak = amplitudes
        self.frange = frange
        self.pk = pk

    def __str__(self):
        return str(self.__dict__)

    __repr__ = __str__


class Frame:
    """5 ms speech/voiced flag record (reference: structs.py:24-34)."""

    def __init__(self, ti, isSpeech, isVoiced):
        self.ti = ti
        self.isSpeech = isSpeech
        self.isVoiced = isVoiced

    def __str__(self):
        return str(self.__dict__)

    __repr__ = __str__
