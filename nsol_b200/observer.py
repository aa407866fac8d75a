"""Observer of the reference API (``nsol.observer.Observer``, nsol/observer.py:18-161):
collects every iterate a solver reports and evaluates measure callbacks afterwards.
Host-side bookkeeping only; a solver with an observer attached copies one iterate
per iteration back from the device."""
import numpy as np


class Observer(object):

    def __init__(self, name="Observer", store_iterates=True):
        """``store_iterates=False`` (additive option): a GPU solver evaluates the measures it can map to
        device reductions (SimilarityMeasures SSD/MAE/MSE/RMSE/PSNR/NCC) per iteration on the device and
        hands over only the final iterate -- no copy of every iterate to the host."""
        self._store_iterates = store_iterates
        self._device_results = None
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
        self._device_results = None

    def get_store_iterates(self):
        return self._store_iterates

    def device_measure_requests(self, n):
        """{name: MeasureRequest} if every measure can be evaluated on the device, else None."""
        from nsol_b200 import _trace
        from nsol_b200.similarity_measures import DEVICE_MEASURES, MeasureRequest
        reqs = {}
        for name, fn in zip(self._measures_names, self._measures):
            try:
                r = fn(_trace.Symbol(("arg",), (int(n),)))
            except Exception:
                return None
            if not isinstance(r, MeasureRequest) or r.kind not in DEVICE_MEASURES:
                return None
            reqs[name] = r
        return reqs

    def set_device_results(self, results):
        """{name: array of one value per iteration (0..n)} produced by a GPU solver."""
        self._device_results = results

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
        if self._device_results is not None:
            for name in self._measures_names:
                self._dic_measures[name] = np.array(self._device_results[name])
            return
        n = len(self._x_list)
        for k, name in enumerate(self._measures_names):
            res = np.zeros(n)
            for i in range(n):
                res[i] = self._measures[k](self._x_list[i])
            self._dic_measures[name] = res
