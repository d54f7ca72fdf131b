//go:build b200 && badger

package option

import (
	"github.com/spf13/pflag"

	"github.com/kubewharf/kubebrain/pkg/storage"
	"github.com/kubewharf/kubebrain/pkg/storage/badger"
)

// engineConfig is the wrapped engine's configuration: exactly what option_badger.go's storageConfig holds
// (cmd/option/option_badger.go:28-50).
type engineConfig struct {
	badger.Config
}

func newEngineConfig() engineConfig {
	e := engineConfig{}
	e.Dir = "./data"
	return e
}

func (e *engineConfig) addFlag(fs *pflag.FlagSet) {
	fs.StringVar(&e.Dir, "data-dir", e.Dir, "data dir of")
}

func (e *engineConfig) validate() error { return nil }

func (e *engineConfig) build() (storage.KvStorage, error) { return badger.NewKvStorage(e.Config) }
