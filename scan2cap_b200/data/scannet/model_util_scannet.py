"""Mirror of the part of data/scannet/model_util_scannet.py:80-172 (ScannetDatasetConfig) that the
detection head and box decoding use: 18 classes, one heading bin, per-class mean box sizes and the
parameter -> oriented-box conversion.  The mean sizes are the 18x3 float64 values of
data/scannet/meta_data/scannet_reference_means.npz (data constants, reproduced digit for digit so box
corners are bit-identical).  The NYU40 label-map helpers need ScanNet's label TSV and belong to the dataset
side (out of scope)."""
import numpy as np

MEAN_SIZE_ARR = np.array([
    [0.7750491029714929, 0.9489772784305719, 0.9654205889420883],
    [1.8690326739217817, 1.8321471223511647, 1.1922299150646347],
    [0.6121477783923587, 0.6192873075057846, 0.7048084833710475],
    [1.4411389838393118, 1.6045203579823017, 0.8365229505964112],
    [1.0478072557954565, 1.2016418836390361, 0.6345700676484581],
    [0.5610123179013166, 0.6084721692226233, 1.7195040055943263],
    [1.0789489470730143, 0.8203399609681988, 1.1692119917347412],
    [0.8417109198057999, 1.3504794475570598, 1.689892503247653],
    [0.2305173710207977, 0.4764049876932717, 0.5656925618884787],
    [1.4548489887322953, 1.9711989456815506, 0.28643280467880305],
    [1.0785803060791836, 1.5370511310202535, 0.8650190604735265],
    [1.4311964378217468, 0.7692311116413818, 1.6498267253793382],
    [0.6296919388045009, 0.7087128690976665, 1.314335867333314],
    [0.4392503422374527, 0.41569593879911637, 1.7000274790657892],
    [0.5850446242623347, 0.5787843832293073, 0.7202961145680844],
    [0.5115869258698381, 0.5096067340403306, 0.3128736034402105],
    [1.1732075942887201, 1.0598714035004377, 0.5181252788752317],
    [0.43294385021345605, 0.5193350711870748, 0.4843745602902239],
], dtype=np.float64)


class ScannetDatasetConfig(object):
    def __init__(self):
        self.type2class = {'cabinet': 0, 'bed': 1, 'chair': 2, 'sofa': 3, 'table': 4, 'door': 5, 'window': 6,
                           'bookshelf': 7, 'picture': 8, 'counter': 9, 'desk': 10, 'curtain': 11,
                           'refrigerator': 12, 'shower curtain': 13, 'toilet': 14, 'sink': 15, 'bathtub': 16,
                           'others': 17}
        self.class2type = {self.type2class[t]: t for t in self.type2class}
        self.mean_size_arr = MEAN_SIZE_ARR.copy()
        self.num_class = len(self.type2class)
        self.num_heading_bin = 1
        self.num_size_cluster = len(self.type2class)
        self.type_mean_size = {self.class2type[i]: self.mean_size_arr[i, :] for i in range(self.num_size_cluster)}

    def class2angle(self, pred_cls, residual, to_label_format=True):
        return 0  # ScanNet boxes are axis-aligned

    def class2angle_batch(self, pred_cls, residual, to_label_format=True):
        return np.zeros(pred_cls.shape[0])

    def size2class(self, size, type_name):
        return self.type2class[type_name], size - self.type_mean_size[type_name]

    def class2size(self, pred_cls, residual):
        return self.mean_size_arr[pred_cls] + residual

    def class2size_batch(self, pred_cls, residual):
        return self.mean_size_arr[pred_cls] + residual

    def param2obb(self, center, heading_class, heading_residual, size_class, size_residual):
        obb = np.zeros((7,))
        obb[0:3] = center
        obb[3:6] = self.class2size(int(size_class), size_residual)
        obb[6] = self.class2angle(heading_class, heading_residual) * -1
        return obb

    def param2obb_batch(self, center, heading_class, heading_residual, size_class, size_residual):
        obb = np.zeros((heading_class.shape[0], 7))
        obb[:, 0:3] = center
        obb[:, 3:6] = self.class2size_batch(size_class, size_residual)
        obb[:, 6] = self.class2angle_batch(heading_class, heading_residual) * -1
        return obb
