"""COCOEvaluator: the evaluation loop around the hot path, same constructor, methods and return values
as yolox/evaluators/coco_evaluator.py:26-217.  The per-batch work runs on the GPU through this package's kernels
(model forward, fused postprocess, detection -> COCO record conversion); only one small D2H copy of the records per batch
reaches the host, instead of the reference's per-image .cpu() + per-detection Python arithmetic.  The final AP numbers
come from the native COCO evaluation (cocoeval.COCOevalBBox) instead of pycocotools."""
import itertools
import json
import time
from typing import List

import torch
import torch.distributed as dist

from . import io as yio
from . import postprocess as pp
from .cocoeval import COCOevalBBox


def _is_main_process() -> bool:
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def _time_synchronized() -> float:
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.time()


def gather_records(records: List[dict], dst: int = 0) -> List[List[dict]]:
    """The reference pickles Python lists through a gloo group (yolox/utils/dist.py:224-265); here the records travel as
    one padded float64 tensor per rank [n, 7] = [image_id, category_id, x, y, w, h, score] in one all-gather."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [records]
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([[r["image_id"], r["category_id"], *r["bbox"], r["score"]] for r in records],
                     dtype=torch.float64, device=dev).reshape(-1, 7)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    cap = max(int(c.item()) for c in counts)
    padded = torch.zeros(max(cap, 1), 7, dtype=torch.float64, device=dev)
    padded[:t.shape[0]] = t
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    if rank != dst:
        return []
    res = []
    for c, o in zip(counts, out):
        rows = o[:int(c.item())].cpu().tolist()
        res.append([dict(image_id=int(r[0]), category_id=int(r[1]), bbox=r[2:6], score=r[6], segmentation=[]) for r in rows])
    return res


class COCOEvaluator:
    """COCO AP evaluation (coco_evaluator.py:26-49: same arguments)."""

    def __init__(self, dataloader, img_size, confthre, nmsthre, num_classes, testdev=False):
        self.dataloader = dataloader
        self.img_size = img_size
        self.confthre = confthre
        self.nmsthre = nmsthre
        self.num_classes = num_classes
        self.testdev = testdev

    # ---- coco_evaluator.py:51-133 ------------------------------------------------------------------------------
    def evaluate(self, model, distributed=False, half=False, trt_file=None, decoder=None, test_size=None):
        """Returns (ap50_95, ap50, summary) on the main process, (0, 0, None) elsewhere.

        half=True deviation (deliberate): the reference keeps the whole post-processing in fp16 --
        torchvision.batched_nms adds fp16 class offsets to fp16 boxes, and `bboxes /= scale`, `obj * cls` run in fp16 on
        the CPU copy (coco_evaluator.py:147-157).  Here the scores / threshold of the selection are evaluated in fp16 like
        the reference, but the IoU tests run in fp32 on the fp16 box values and the record arithmetic (unscale, xywh,
        score product) runs in fp32: records and AP under half=True are therefore closer to the fp32 evaluation than the
        reference's, not bit-identical to it.  With half=False everything is fp32 on both sides and the records agree."""
        if trt_file is not None:
            raise NotImplementedError("TensorRT engines (torch2trt) are outside this package; pass the model itself")
        model = model.eval()
        if half:
            model = model.half()
        dev = next(model.parameters()).device
        dtype = torch.float16 if half else torch.float32
        data_list = []
        inference_time = 0.0
        nms_time = 0.0
        n_samples = max(len(self.dataloader) - 1, 1)
        for cur_iter, (imgs, _, info_imgs, ids) in enumerate(self.dataloader):
            with torch.no_grad():
                imgs = imgs.to(dev, dtype, non_blocking=True)
                is_time_record = cur_iter < len(self.dataloader) - 1   # the last batch may be short (:104-105)
                if is_time_record:
                    start = time.time()
                outputs = model(imgs)
                if decoder is not None:
                    outputs = decoder(outputs, dtype=outputs.type())
                if is_time_record:
                    infer_end = _time_synchronized()
                    inference_time += infer_end - start
                det, cnt, _ = pp.postprocess_raw(outputs, self.num_classes, self.confthre, self.nmsthre)
                if is_time_record:
                    nms_end = _time_synchronized()
                    nms_time += nms_end - infer_end
            data_list.extend(self._records_dense(det, cnt, info_imgs, ids))
        statistics = torch.tensor([inference_time, nms_time, n_samples], dtype=torch.float32, device=dev)
        if distributed:
            data_list = gather_records(data_list, dst=0)
            data_list = list(itertools.chain(*data_list))
            dist.reduce(statistics, dst=0)
        eval_results = self.evaluate_prediction(data_list, statistics)
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
        return eval_results

    def _class_ids(self):
        return list(self.dataloader.dataset.class_ids)

    def _records_dense(self, det, cnt, info_imgs, ids) -> List[dict]:
        """Device-side convert_to_coco_format for the fixed-shape postprocess result: one kernel, one D2H copy."""
        hw = [(float(h), float(w)) for h, w in zip(info_imgs[0], info_imgs[1])]
        rec = yio.coco_records(det, cnt, hw, self.img_size, self._class_ids())
        rec_h, cnt_h = rec.cpu(), cnt.cpu().tolist()
        out = []
        for b, (n, img_id) in enumerate(zip(cnt_h, ids)):
            rows = rec_h[b, :n].tolist()
            out.extend(dict(image_id=int(img_id), category_id=int(r[5]), bbox=r[0:4], score=r[4], segmentation=[])
                       for r in rows)
        return out

    # ---- coco_evaluator.py:135-165: same signature, for callers that hold the reference's list of [n,7] tensors ---
    def convert_to_coco_format(self, outputs, info_imgs, ids):
        data_list = []
        class_ids = self._class_ids()
        for (output, img_h, img_w, img_id) in zip(outputs, info_imgs[0], info_imgs[1], ids):
            if output is None:
                continue
            output = output.cpu().float()
            bboxes = output[:, 0:4]
            scale = min(self.img_size[0] / float(img_h), self.img_size[1] / float(img_w))
            bboxes /= scale
            bboxes[:, 2] = bboxes[:, 2] - bboxes[:, 0]          # xyxy2xywh (yolox/utils/boxes.py)
            bboxes[:, 3] = bboxes[:, 3] - bboxes[:, 1]
            cls = output[:, 6]
            scores = output[:, 4] * output[:, 5]
            for ind in range(bboxes.shape[0]):
                data_list.append({"image_id": int(img_id), "category_id": class_ids[int(cls[ind])],
                                  "bbox": bboxes[ind].numpy().tolist(), "score": scores[ind].numpy().item(),
                                  "segmentation": []})
        return data_list

    # ---- coco_evaluator.py:167-217 ---------------------------------------------------------------------------
    def evaluate_prediction(self, data_dict, statistics):
        if not _is_main_process():
            return 0, 0, None
        inference_time = statistics[0].item()
        nms_time = statistics[1].item()
        n_samples = statistics[2].item()
        bs = self.dataloader.batch_size
        a_infer_time = 1000 * inference_time / (n_samples * bs)
        a_nms_time = 1000 * nms_time / (n_samples * bs)
        time_info = ", ".join("Average {} time: {:.2f} ms".format(k, v) for k, v in
                              zip(["forward", "NMS", "inference"], [a_infer_time, a_nms_time, a_infer_time + a_nms_time]))
        info = time_info + "\n"
        if len(data_dict) > 0:
            if self.testdev:
                json.dump(data_dict, open("./yolox_testdev_2017.json", "w"))
            ev = COCOevalBBox(self.dataloader.dataset.coco, data_dict).evaluate()
            info += ev.summarize()
            return ev.stats[0], ev.stats[1], info
        return 0, 0, info
