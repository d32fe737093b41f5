"""Value types and the abstract plugin interface of the m3vision pipeline that AePoseEstimator plugs into.
Mirrors the ABCs vendored at auto_pose/m3_interface/m3_interfaces.py:57-211 (PoseEstimate, PoseEstInterface,
BoundingBox) -- only what the pose-estimation plugin surface needs."""
from abc import ABCMeta, abstractmethod

import numpy as np


class PoseEstimate(object):
    def __init__(self, name='SLC', trafo=np.identity(4), quality=1.0):
        self.name = name
        self.trafo = trafo
        self.quality = quality


class BoundingBox(object):
    """Normalised [0,1] box with a {class: score} dict."""

    def __init__(self, xmin=0.0, ymin=0.0, xmax=1.0, ymax=1.0, classes=None):
        self.xmin, self.ymin, self.xmax, self.ymax = xmin, ymin, xmax, ymax
        self.classes = classes if classes is not None else {'SLC': 1.0}


class PoseEstInterface(metaclass=ABCMeta):

    def __init__(self, configpath=None, m3vision_cfg=None):
        pass

    @abstractmethod
    def set_parameter(self, string_name, string_val):
        pass

    def get_params(self, config):
        """str path (.yml/.yaml or INI) or an already parsed object (m3_interfaces.py:99-119)."""
        if isinstance(config, str):
            if '.yml' in config or '.yaml' in config:
                import yaml
                with open(config, 'r') as f:
                    return yaml.safe_load(f)
            import configparser
            params = configparser.ConfigParser(inline_comment_prefixes="#")
            params.read(config)
            return params
        return config

    @abstractmethod
    def query_process_requirements(self):
        return ['color_img', 'depth_img', 'camK', 'camPose']

    @abstractmethod
    def query_image_format(self):
        return {'color_format': 'rgb', 'color_data_type': np.float32, 'depth_data_type': np.float32}

    @abstractmethod
    def process(self, bboxes=[], color_img=None, depth_img=None, camK=None, camPose=None, rois3ds=[]):
        pass
