/*
 * harness_prelude.js — TEST INFRASTRUCTURE.  Environment for running the
 * reference's lib/register.js, unmodified, on the ES3 engine of the reference
 * tree: a `require` that serves small stand-ins for the npm modules that are
 * not vendored (assert-plus, once, vasync) and for the two Node core modules
 * the file uses (os, path).  Nothing here restates register.js itself.
 *
 * `path` follows the documented algorithm of Node's posix path.join /
 * normalize / dirname (Node core is not in the reference tree).
 */
var module = { exports: {} };
var __hostname = '';
var __timers = [];

function setTimeout(fn, ms) { __timers.push(fn); return (__timers.length); }
function __runTimers() { while (__timers.length) { (__timers.shift())(); } }

if (!Function.prototype.bind) {
    Function.prototype.bind = function (self) {
        var fn = this;
        var pre = Array.prototype.slice.call(arguments, 1);
        return (function () {
            return (fn.apply(self, pre.concat(Array.prototype.slice.call(arguments))));
        });
    };
}
if (!Object.keys) {
    Object.keys = function (o) { var k = [], p; for (p in o) { if (o.hasOwnProperty(p)) k.push(p); } return (k); };
}
if (!Array.isArray) {
    Array.isArray = function (a) { return (Object.prototype.toString.call(a) === '[object Array]'); };
}

var __path = (function () {
    function normalizeSegments(parts, allowAboveRoot) {
        var res = [], i, p;
        for (i = 0; i < parts.length; i++) {
            p = parts[i];
            if (!p || p === '.')
                continue;
            if (p === '..') {
                if (res.length && res[res.length - 1] !== '..')
                    res.pop();
                else if (allowAboveRoot)
                    res.push('..');
            } else {
                res.push(p);
            }
        }
        return (res);
    }
    function normalize(p) {
        var isAbsolute = p.charAt(0) === '/';
        var trailingSlash = p.length > 0 && p.charAt(p.length - 1) === '/';
        p = normalizeSegments(p.split('/'), !isAbsolute).join('/');
        if (!p && !isAbsolute)
            p = '.';
        if (p && trailingSlash)
            p += '/';
        return ((isAbsolute ? '/' : '') + p);
    }
    function join() {
        var parts = [], i;
        for (i = 0; i < arguments.length; i++) {
            if (typeof (arguments[i]) !== 'string')
                throw new TypeError('Arguments to path.join must be strings');
            if (arguments[i])
                parts.push(arguments[i]);
        }
        return (normalize(parts.join('/')));
    }
    function dirname(p) {
        if (p.length === 0)
            return ('.');
        var hasRoot = p.charAt(0) === '/', end = -1, matchedSlash = true, i;
        for (i = p.length - 1; i >= 1; --i) {
            if (p.charAt(i) === '/') {
                if (!matchedSlash) { end = i; break; }
            } else {
                matchedSlash = false;
            }
        }
        if (end === -1)
            return (hasRoot ? '/' : '.');
        if (hasRoot && end === 1)
            return ('//');
        return (p.slice(0, end));
    }
    return ({ join: join, normalize: normalize, dirname: dirname });
})();

var __assert = (function () {
    function fail(name, type) { throw new Error(name + ' (' + type + ') is required'); }
    function isArr(a) { return (Array.isArray(a)); }
    var a = {
        ok: function (v, msg) { if (!v) throw new Error(msg || 'assertion failed'); },
        string: function (v, n) { if (typeof (v) !== 'string') fail(n, 'string'); },
        number: function (v, n) { if (typeof (v) !== 'number') fail(n, 'number'); },
        func: function (v, n) { if (typeof (v) !== 'function') fail(n, 'func'); },
        object: function (v, n) { if (typeof (v) !== 'object' || v === null) fail(n, 'object'); },
        arrayOfString: function (v, n) {
            var i;
            if (!isArr(v)) fail(n, '[string]');
            for (i = 0; i < v.length; i++) if (typeof (v[i]) !== 'string') fail(n, '[string]');
        },
        arrayOfNumber: function (v, n) {
            var i;
            if (!isArr(v)) fail(n, '[number]');
            for (i = 0; i < v.length; i++) if (typeof (v[i]) !== 'number') fail(n, '[number]');
        },
        arrayOfObject: function (v, n) {
            var i;
            if (!isArr(v)) fail(n, '[object]');
            for (i = 0; i < v.length; i++) if (typeof (v[i]) !== 'object') fail(n, '[object]');
        }
    };
    function optional(f) { return (function (v, n) { if (v !== undefined) f(v, n); }); }
    a.optionalString = optional(a.string);
    a.optionalNumber = optional(a.number);
    a.optionalObject = optional(a.object);
    a.optionalArrayOfNumber = optional(a.arrayOfNumber);
    a.optionalArrayOfString = optional(a.arrayOfString);
    return (a);
})();

function __once(fn) {
    var called = false, value;
    return (function () {
        if (called)
            return (value);
        called = true;
        value = fn.apply(this, arguments);
        return (value);
    });
}

var __vasync = {
    forEachParallel: function (args, cb) {
        var inputs = args.inputs, pending = inputs.length, firstErr = null, done = false, i;
        if (pending === 0) { cb(null, { operations: [] }); return; }
        function mk() {
            var fired = false;
            return (function (err) {
                if (fired) return;
                fired = true;
                if (err && !firstErr) firstErr = err;
                if (--pending === 0 && !done) { done = true; cb(firstErr, { operations: [] }); }
            });
        }
        for (i = 0; i < inputs.length; i++)
            args.func(inputs[i], mk());
    },
    forEachPipeline: function (args, cb) {
        var inputs = args.inputs, i = 0;
        function next(err) {
            if (err) { cb(err); return; }
            if (i >= inputs.length) { cb(null); return; }
            args.func(inputs[i++], next);
        }
        next();
    },
    pipeline: function (args, cb) {
        var funcs = args.funcs, i = 0;
        function next(err) {
            if (err) { cb(err); return; }
            if (i >= funcs.length) { cb(null); return; }
            funcs[i++](args.arg, next);
        }
        next();
    }
};

function require(name) {
    switch (name) {
    case 'os':
        return ({
            hostname: function () { return (__hostname); },
            networkInterfaces: function () {
                return ({ lo0: [ { address: '127.0.0.1', internal: true } ],
                    net0: [ { address: '10.77.77.7', internal: false } ] });
            }
        });
    case 'path': return (__path);
    case 'assert-plus': return (__assert);
    case 'once': return (__once);
    case 'vasync': return (__vasync);
    default: throw new Error('module not available in the harness: ' + name);
    }
}
