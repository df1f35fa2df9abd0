"""Entry point: `python -m src.main <config.yaml> [--test]` (reference src/main.py:19-190).

The YAML config is turned into objects purely by name lookup - `getattr(module, section.name)(**section.kwargs)` -
over the same modules as the reference: src.data.datasets, src.data.dataloader, src.model.nets, torch.nn /
src.model.losses, src.model.metrics, torch.optim (+ pvsr.optim), torch.optim.lr_scheduler, src.callbacks.loggers,
src.callbacks.monitor, src.runner.trainers / predictors.  Reference configs load unchanged (only data paths differ).

Launched under torchrun (WORLD_SIZE > 1) it becomes data-parallel training / sequence-sharded testing: every rank
uses the GPU LOCAL_RANK, dataloaders are sharded, rank 0 logs and saves.
"""
import argparse
import copy
import logging
import random
from pathlib import Path

import torch
import yaml

import src
from pvsr import optim as pvsr_optim
from pvsr import parallel


class Config(dict):
    """Attribute-access dict with the handful of `Box` methods the entry point relies on
    (python-box is not installed here): from_yaml, to_dict, and dict's own get / pop / update."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in list(self.items()):
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, Config):
            return cls(v)
        if isinstance(v, list):
            return [cls._wrap(x) for x in v]
        return v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = self._wrap(value)

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = self._wrap(v)

    @classmethod
    def from_yaml(cls, filename):
        with open(filename) as f:
            return cls(yaml.safe_load(f))

    def to_dict(self):
        def plain(v):
            if isinstance(v, dict):
                return {k: plain(x) for k, x in v.items()}
            if isinstance(v, list):
                return [plain(x) for x in v]
            return str(v) if isinstance(v, Path) else v
        return plain(self)


def _get_instance(module, config, *args):
    """`module.<config.name>(*args, **config.kwargs)`."""
    cls = getattr(module, config.name)
    kwargs = config.get('kwargs')
    return cls(*args, **kwargs) if kwargs else cls(*args)


def _build_losses(config):
    loss_fns, loss_weights = [], []
    for entry in config.losses:
        module = torch.nn if 'Loss' in entry.name and hasattr(torch.nn, entry.name) else src.model.losses
        loss_fns.append(_get_instance(module, entry))
        loss_weights.append(entry.weight)
    return loss_fns, loss_weights


def _device(name, local_rank, world):
    if 'cuda' in name and not torch.cuda.is_available():
        raise ValueError("The cuda is not available. The B200 RefineNet path has no CPU fallback.")
    return torch.device('cuda', local_rank) if world > 1 and 'cuda' in name else torch.device(name)


def main(args):
    logging.info(f'Load the config from "{args.config_path}".')
    config = Config.from_yaml(args.config_path)
    rank, local_rank, world = parallel.env_world()
    saved_dir = Path(config.main.saved_dir)
    if rank == 0:
        saved_dir.mkdir(parents=True, exist_ok=True)
        with open(saved_dir / 'config.yaml', 'w+') as f:
            yaml.dump(config.to_dict(), f, default_flow_style=False)

    section = config.predictor if args.test else config.trainer
    device = _device(section.kwargs.device, local_rank, world)
    parallel.init(device=device if device.type == 'cuda' else None)
    data_dir = config.dataset.kwargs.get('data_dir')
    data_dir = Path(data_dir) if data_dir is not None else None
    cls = getattr(src.data.datasets, config.dataset.name)
    collate_fn = getattr(cls, 'collate_fn', None)

    if not args.test:
        # deterministic experiment: the reference seeds python's RNG with the config string and torch from it
        random.seed(config.main.random_seed)
        seed = random.getstate()[1][1]
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)

        datasets = {}
        for kind in ('train', 'valid'):
            config.dataset.kwargs.update(data_dir=data_dir, type=kind)
            datasets[kind] = _get_instance(src.data.datasets, config.dataset)
        loader_kw = copy.copy(config.dataloader.kwargs)
        train_bs, valid_bs = loader_kw.pop('train_batch_size'), loader_kw.pop('valid_batch_size')
        loader_cfg = Config(name=config.dataloader.name, kwargs=dict(loader_kw, collate_fn=collate_fn,
                                                                      shard=(rank, world)))
        loader_cfg.kwargs.update(batch_size=train_bs)
        train_dataloader = _get_instance(src.data.dataloader, loader_cfg, datasets['train'])
        loader_cfg.kwargs.update(batch_size=valid_bs, shard_pad=False)   # under DDP: each validation sample counted once
        valid_dataloader = _get_instance(src.data.dataloader, loader_cfg, datasets['valid'])

        net = _get_instance(src.model.nets, config.net)
        loss_fns, loss_weights = _build_losses(config)
        metric_fns = [_get_instance(src.model.metrics, m) for m in config.metrics]

        if hasattr(pvsr_optim, config.optimizer.name):
            # flat-buffer optimisers need the parameters on their final device first
            net = net.to(device)
            optimizer = getattr(pvsr_optim, config.optimizer.name).for_net(net, **(config.optimizer.get('kwargs') or {}))
        else:
            optimizer = _get_instance(torch.optim, config.optimizer, net.parameters())
        lr_scheduler = (_get_instance(torch.optim.lr_scheduler, config.lr_scheduler, optimizer)
                        if config.get('lr_scheduler') else None)

        logger = None
        if rank == 0:
            dummy = torch.randn(tuple(config.logger.kwargs.dummy_input))
            config.logger.kwargs.update(log_dir=saved_dir / 'log', net=net, dummy_input=dummy)
            logger = _get_instance(src.callbacks.loggers, config.logger)
        config.monitor.kwargs.update(checkpoints_dir=saved_dir / 'checkpoints')
        monitor = _get_instance(src.callbacks.monitor, config.monitor)

        config.trainer.kwargs.update(device=device, train_dataloader=train_dataloader,
                                     valid_dataloader=valid_dataloader, net=net, loss_fns=loss_fns,
                                     loss_weights=loss_weights, metric_fns=metric_fns, optimizer=optimizer,
                                     lr_scheduler=lr_scheduler, logger=logger, monitor=monitor)
        trainer = _get_instance(src.runner.trainers, config.trainer)
        loaded_path = config.main.get('loaded_path')
        if loaded_path:
            logging.info(f'Load the previous checkpoint from "{loaded_path}".')
            trainer.load(Path(loaded_path))
        logging.info('Resume training.' if loaded_path else 'Start training.')
        trainer.train()
        logging.info('End training.')
    else:
        config.dataset.kwargs.update(data_dir=data_dir, type='test')
        test_dataset = _get_instance(src.data.datasets, config.dataset)
        loader_cfg = Config(name=config.dataloader.name, kwargs=dict(config.dataloader.kwargs, collate_fn=collate_fn))
        test_dataloader = _get_instance(src.data.dataloader, loader_cfg, test_dataset)
        net = _get_instance(src.model.nets, config.net)
        loss_fns, loss_weights = _build_losses(config)
        metric_fns = [_get_instance(src.model.metrics, m) for m in config.metrics]
        config.predictor.kwargs.update(device=device, test_dataloader=test_dataloader, net=net, loss_fns=loss_fns,
                                       loss_weights=loss_weights, metric_fns=metric_fns)
        predictor = _get_instance(src.runner.predictors, config.predictor)
        if config.net.name != 'Bicubic' and config.main.get('loaded_path'):
            logging.info(f'Load the previous checkpoint from "{config.main.loaded_path}".')
            predictor.load(Path(config.main.loaded_path))
        logging.info('Start testing.')
        predictor.predict()
        logging.info('End testing.')


def _parse_args():
    parser = argparse.ArgumentParser(description='The script for the training and the testing.')
    parser.add_argument('config_path', type=Path, help='The path of the config file.')
    parser.add_argument('--test', action='store_true', help='Perform the testing if specified; otherwise the training.')
    return parser.parse_args()


if __name__ == '__main__':
    logging.basicConfig(format='%(asctime)s | %(levelname)s | %(message)s', level=logging.INFO,
                        datefmt='%Y-%m-%d %H:%M:%S')
    main(_parse_args())
