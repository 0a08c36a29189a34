"""ORACLE shim: marker base classes (pybasicbayes.abstractions)."""


class GibbsSampling(object):
    def resample(self, data=[]):
        raise NotImplementedError


class ModelGibbsSampling(object):
    def resample_model(self):
        raise NotImplementedError
