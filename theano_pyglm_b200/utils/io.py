"""Data / model / results files (interface of pyglm/utils/io.py:82-149 and the pickles written by
test/generate_synth_data.py:101-135 and test/synth_map.py).

The on-disk format is the reference's: `data.pkl` = the data dict {'S','N','dt','T','stim','dt_stim','vars'?,
'X'?}, `model.pkl` = the model dict, `results.pkl` = a state dict or a list of them.  Files written by the
Python-2 reference load with encoding='latin1'.  Engine handles (`_b200`) never reach the disk."""
import copy
import os
import pickle

import numpy as np


def _strip(obj):
    """Drop device handles / derived arrays that must not be pickled."""
    if isinstance(obj, dict):
        return {k: _strip(v) for k, v in obj.items() if k not in ('_b200',)}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_strip(v) for v in obj)
    return obj


def save_pickle(obj, path):
    with open(path, 'wb') as f:
        pickle.dump(_strip(obj), f, protocol=2)             # protocol 2: readable from the reference's cPickle


def load_pickle(path):
    with open(path, 'rb') as f:
        try:
            return pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            return pickle.load(f, encoding='latin1')


def load_data(data_file, verbose=False):
    """io.py:82-125: a .pkl data dict or a .mat file with the same fields."""
    if data_file is None:
        raise Exception("Path to data file (.mat or .pkl) must be specified. "
                        "To generate synthetic data use theano_pyglm_b200.utils.synth.")
    if data_file.endswith('.mat'):
        import scipy.io
        data = scipy.io.loadmat(data_file, squeeze_me=True)
        data['N'] = int(data['N'])
        data['T'] = float(data['T'])
    elif data_file.endswith('.pkl'):
        data = load_pickle(data_file)
        if verbose:
            print("Data has %d neurons, %d spikes, and %d time bins at %.3fHz sample rate"
                  % (data['N'], np.sum(data['S']), data['S'].shape[0], 1.0 / data['dt']))
    else:
        raise Exception("Unrecognized file type: %s" % data_file)
    data.pop('preprocessed', None)
    return data


def segment_data(data, T_range):
    """io.py:127-151: the sub-recording [T_start, T_stop) seconds (spikes and stimulus)."""
    T_start, T_stop = T_range
    assert 0 <= T_start <= data['T'] and 0 <= T_stop <= data['T'] and T_start < T_stop
    new = copy.deepcopy(_strip(data))
    new.pop('preprocessed', None)
    new.pop('fstim', None)
    new['T'] = T_stop - T_start
    i0, i1 = int(T_start // data['dt']), int(T_stop // data['dt'])
    new['S'] = new['S'][i0:i1, :]
    if new.get('stim') is not None:
        j0, j1 = int(T_start // data['dt_stim']), int(T_stop // data['dt_stim'])
        new['stim'] = new['stim'][j0:j1, :]
    return new


def save_results(results_dir, data=None, model=None, results=None):
    """data.pkl / model.pkl / results.pkl in `results_dir` (generate_synth_data.py:101-135, synth_map.py)."""
    os.makedirs(results_dir, exist_ok=True)
    for name, obj in (('data.pkl', data), ('model.pkl', model), ('results.pkl', results)):
        if obj is not None:
            save_pickle(obj, os.path.join(results_dir, name))
