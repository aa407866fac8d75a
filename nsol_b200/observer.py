"""Observer of the reference API (``nsol.observer.Observer``, nsol/observer.py:18-161):
collects every iterate a solver reports and evaluates measure callbacks afterwards.
Host-side bookkeeping only; a solver with an observer attached copies one iterate
per iteration back from the device."""
import numpy as np


class Observer(object):

    def __init__(self, name="Observer"):
        self._name = name
        self._x_list = []
        self._measures = []
        self._measures_names = []
        self._dic_measures = {}
        self._computational_time = None

    def add_x(self, x):
        self._x_list.append(x)

    def set_name(self, name):
        self._name = name

    def get_name(self):
        return self._name

    def clear_x_list(self):
        self._x_list = []

    def get_x_list(self):
        return self._x_list

    def set_measures(self, measures_dic):
        for name in list(measures_dic.keys()):
            self._measures_names.append(name)
            self._measures.append(measures_dic[name])
            self._dic_measures.update({name: None})

    def get_measures(self):
        return self._dic_measures

    def set_computational_time(self, computational_time):
        self._computational_time = computational_time

    def get_computational_time(self):
        return self._computational_time

    def compute_measures(self):
        """One value per stored iterate and measure (nsol/observer.py:111-119)."""
        n = len(self._x_list)
        for k, name in enumerate(self._measures_names):
            res = np.zeros(n)
            for i in range(n):
                res[i] = self._measures[k](self._x_list[i])
            self._dic_measures[name] = res
