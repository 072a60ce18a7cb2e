"""scikit-garden_b200: the quantile-forest predict path of scikit-garden, B200-native.

Import as ``skgarden_b200`` (the shim ``skgarden_b200.py`` at the repo root maps the
name to this directory).  Public surface = the reference's for this path
(skgarden/quantile/__init__.py:1-10):

    RandomForestQuantileRegressor, ExtraTreesQuantileRegressor,
    DecisionTreeQuantileRegressor, ExtraTreeQuantileRegressor
    RandomForestRegressor, ExtraTreesRegressor      (skgarden/forest.py: predict(X, return_std=True))

plus the lower-level pieces the estimators are made of: ``flatten_forest`` /
``FlatForest`` (host layout), ``DeviceForest`` (HBM-resident forest + kernel calls).
"""
from .ensemble import ExtraTreesQuantileRegressor, RandomForestQuantileRegressor
from .flatten import FlatForest, bootstrap_counts, flatten_forest, flatten_forest_device
from .tree import DecisionTreeQuantileRegressor, ExtraTreeQuantileRegressor
from .forest import ExtraTreesRegressor, RandomForestRegressor

__all__ = ["RandomForestQuantileRegressor", "ExtraTreesQuantileRegressor",
           "DecisionTreeQuantileRegressor", "ExtraTreeQuantileRegressor",
           "RandomForestRegressor", "ExtraTreesRegressor", "FlatForest",
           "flatten_forest", "flatten_forest_device", "bootstrap_counts", "DeviceForest"]


def __getattr__(name):          # DeviceForest pulls in torch; load it on first use
    if name == "DeviceForest":
        from .device_forest import DeviceForest
        return DeviceForest
    raise AttributeError(name)
